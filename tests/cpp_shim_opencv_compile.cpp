// compile-and-link check of the LITIV_B200_WITH_OPENCV branch of the header-only C++ drop-in: the classes derive from
// cv::BackgroundSubtractor and take cv::Mat / cv::InputArray / cv::OutputArray like the reference's. No OpenCV C++ exists in this image, so
// the test compiles against oracle/cvcompat (the minimal OpenCV-compatible declarations written to build the reference itself); with a
// real OpenCV the same header is used unchanged. No GPU needed: only the error paths run.
#define LITIV_B200_WITH_OPENCV
#include "litiv_b200.hpp"
#include <cstdio>
#include <cstring>
int main() {
    cv::Mat img(48, 64, CV_8UC3, cv::Scalar(7, 7, 7)), mask, bg, roi(48, 64, CV_8UC1, cv::Scalar(255));
    try {
        lvb::BackgroundSubtractorSuBSENSE s;                    // throws without a device: "no CPU fallback"
        cv::BackgroundSubtractor& base = s;                     // drop-in: usable through the OpenCV interface
        s.validateROI(roi);
        s.initialize(img, roi);
        base.apply(img, mask, 1.0);
        base.getBackgroundImage(bg);
        s.setROI(roi);
        cv::Mat r = s.getROICopyMat();
        std::printf("ran on GPU: mask %dx%d type %d, bg type %d, roi corner %d\n", mask.cols, mask.rows, mask.type(), bg.type(), (int)r.data[0]);
    } catch(const lvb::Exception& e) { std::printf("lvb::Exception: %s\n", e.what()); }
    {   // validateROI is pure host code: the 2-px border is cleared, the rest untouched
        cv::Mat r2(9, 11, CV_8UC1, cv::Scalar(255));
        uint8_t buf[9 * 11]; std::memset(buf, 255, sizeof(buf));
        lvb_validate_roi(buf, 11, 9, 2);
        int nz = 0; for(uint8_t v : buf) nz += v != 0;
        std::printf("validateROI keeps %d of 99\n", nz);
    }
    try { lvb::BackgroundSubtractorLOBSTER l; cv::Mat g(48, 64, CV_8UC1, cv::Scalar(3)); l.initialize(g, cv::Mat()); l.apply(g, mask, 16.0); } catch(const lvb::Exception& e) { std::printf("lvb::Exception: %s\n", e.what()); }
    try { lvb::BackgroundSubtractorViBe_3ch v; v.initialize(lvb::ImageView(img)); v.apply(img, mask, 16.0); v.getBackgroundImage(bg); } catch(const lvb::Exception& e) { std::printf("lvb::Exception: %s\n", e.what()); }
    return 0;
}
