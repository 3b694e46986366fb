// ORACLE — TEST INFRASTRUCTURE ONLY. cvcompat: cv::BackgroundSubtractor (the interface IIBackgroundSubtractor derives from).
#pragma once
#include "../core.hpp"
namespace cv {
class BackgroundSubtractor : public Algorithm {
public:
    virtual void apply(InputArray image, OutputArray fgmask, double learningRate = -1) = 0;
    virtual void getBackgroundImage(OutputArray backgroundImage) const = 0;
};
} // namespace cv
