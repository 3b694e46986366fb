// ORACLE — TEST INFRASTRUCTURE ONLY. cvcompat: the imgproc calls the reference's hot path makes (see core.hpp header comment).
#pragma once
#include "core.hpp"

namespace cv {

enum InterpolationFlags { INTER_NEAREST = 0, INTER_LINEAR = 1, INTER_CUBIC = 2, INTER_AREA = 3, INTER_LANCZOS4 = 4 };
enum MorphTypes { MORPH_ERODE = 0, MORPH_DILATE = 1, MORPH_OPEN = 2, MORPH_CLOSE = 3, MORPH_GRADIENT = 4, MORPH_TOPHAT = 5, MORPH_BLACKHAT = 6 };
enum MorphShapes { MORPH_RECT = 0, MORPH_CROSS = 1, MORPH_ELLIPSE = 2 };
enum ColorConversionCodes { COLOR_BGR2GRAY = 6, COLOR_GRAY2BGR = 8, COLOR_BGR2YCrCb = 36, COLOR_YCrCb2BGR = 38, COLOR_BGR2HSV = 40, COLOR_BGR2Lab = 44 };
enum LineTypes { FILLED = -1, LINE_4 = 4, LINE_8 = 8, LINE_AA = 16 };
enum HersheyFonts { FONT_HERSHEY_SIMPLEX = 0, FONT_HERSHEY_PLAIN = 1, FONT_HERSHEY_DUPLEX = 2 };
enum FloodFillFlags { FLOODFILL_FIXED_RANGE = 1 << 16, FLOODFILL_MASK_ONLY = 1 << 17 };

inline Scalar morphologyDefaultBorderValue() { return Scalar::all(1.7976931348623157e308); }
Mat getStructuringElement(int shape, Size ksize, Point anchor = Point(-1, -1));
void resize(InputArray src, OutputArray dst, Size dsize, double fx = 0, double fy = 0, int interpolation = INTER_LINEAR);
void medianBlur(InputArray src, OutputArray dst, int ksize);
void blur(InputArray src, OutputArray dst, Size ksize, Point anchor = Point(-1, -1), int borderType = BORDER_DEFAULT);
void GaussianBlur(InputArray src, OutputArray dst, Size ksize, double sigmaX, double sigmaY = 0, int borderType = BORDER_DEFAULT);
void erode(InputArray src, OutputArray dst, InputArray kernel, Point anchor = Point(-1, -1), int iterations = 1, int borderType = BORDER_CONSTANT,
           const Scalar& borderValue = morphologyDefaultBorderValue());
void dilate(InputArray src, OutputArray dst, InputArray kernel, Point anchor = Point(-1, -1), int iterations = 1, int borderType = BORDER_CONSTANT,
            const Scalar& borderValue = morphologyDefaultBorderValue());
void morphologyEx(InputArray src, OutputArray dst, int op, InputArray kernel, Point anchor = Point(-1, -1), int iterations = 1,
                  int borderType = BORDER_CONSTANT, const Scalar& borderValue = morphologyDefaultBorderValue());
int floodFill(InputOutputArray image, Point seedPoint, Scalar newVal, Rect* rect = nullptr, Scalar loDiff = Scalar(), Scalar upDiff = Scalar(), int flags = 4);
void accumulateWeighted(InputArray src, InputOutputArray dst, double alpha, InputArray mask = noArray());
void accumulateProduct(InputArray src1, InputArray src2, InputOutputArray dst, InputArray mask = noArray());
void cvtColor(InputArray src, OutputArray dst, int code, int dstCn = 0);
void Scharr(InputArray src, OutputArray dst, int ddepth, int dx, int dy, double scale = 1, double delta = 0, int borderType = BORDER_DEFAULT);
// drawing (debug displays of the reference only; no-ops here)
void circle(InputOutputArray img, Point center, int radius, const Scalar& color, int thickness = 1, int lineType = LINE_8, int shift = 0);
void putText(InputOutputArray img, const String& text, Point org, int fontFace, double fontScale, Scalar color, int thickness = 1, int lineType = LINE_8,
             bool bottomLeftOrigin = false);
void rectangle(InputOutputArray img, Rect rec, const Scalar& color, int thickness = 1, int lineType = LINE_8, int shift = 0);
void line(InputOutputArray img, Point pt1, Point pt2, const Scalar& color, int thickness = 1, int lineType = LINE_8, int shift = 0);

} // namespace cv
