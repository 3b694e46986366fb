// ORACLE — TEST INFRASTRUCTURE ONLY. cvcompat: cv::Feature2D base (LBSP derives from cv::DescriptorExtractor) and the key-point filter.
#pragma once
#include "core.hpp"
namespace cv {
class KeyPointsFilter {
public:
    static void runByImageBorder(std::vector<KeyPoint>& keypoints, Size imageSize, int borderSize);
    static void runByPixelsMask(std::vector<KeyPoint>& keypoints, const Mat& mask);
};
class Feature2D : public virtual Algorithm {
public:
    virtual ~Feature2D() {}
    virtual void detect(InputArray image, std::vector<KeyPoint>& keypoints, InputArray mask = noArray());
    virtual void detect(InputArrayOfArrays images, std::vector<std::vector<KeyPoint>>& keypoints, InputArrayOfArrays masks = noArray());
    virtual void compute(InputArray image, std::vector<KeyPoint>& keypoints, OutputArray descriptors);
    virtual void compute(InputArrayOfArrays images, std::vector<std::vector<KeyPoint>>& keypoints, OutputArrayOfArrays descriptors);
    virtual void detectAndCompute(InputArray image, InputArray mask, std::vector<KeyPoint>& keypoints, OutputArray descriptors, bool useProvidedKeypoints = false);
    virtual int descriptorSize() const { return 0; }
    virtual int descriptorType() const { return CV_32F; }
    virtual int defaultNorm() const { return NORM_L2; }
    virtual bool empty() const override { return true; }
    virtual void read(const FileNode&) override {}
    virtual void write(FileStorage&) const override {}
    virtual String getDefaultName() const override { return "Feature2D"; }
};
typedef Feature2D FeatureDetector;
typedef Feature2D DescriptorExtractor;
} // namespace cv
