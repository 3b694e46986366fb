// ORACLE — TEST INFRASTRUCTURE ONLY. cvcompat: highgui names used by the reference's debug displays (no-ops).
#pragma once
#include "core.hpp"
namespace cv {
enum WindowFlags { WINDOW_NORMAL = 0, WINDOW_AUTOSIZE = 1 };
void imshow(const String& winname, InputArray mat);
int waitKey(int delay = 0);
void namedWindow(const String& winname, int flags = WINDOW_AUTOSIZE);
void destroyWindow(const String& winname);
void destroyAllWindows();
void moveWindow(const String& winname, int x, int y);
void resizeWindow(const String& winname, int width, int height);
typedef void (*MouseCallback)(int event, int x, int y, int flags, void* userdata);
void setMouseCallback(const String& winname, MouseCallback onMouse, void* userdata = nullptr);
bool imwrite(const String& filename, InputArray img, const std::vector<int>& params = std::vector<int>());
Mat imread(const String& filename, int flags = 1);
enum MouseEventTypes { EVENT_MOUSEMOVE = 0, EVENT_LBUTTONDOWN = 1, EVENT_RBUTTONDOWN = 2, EVENT_MBUTTONDOWN = 3, EVENT_LBUTTONUP = 4, EVENT_RBUTTONUP = 5, EVENT_MBUTTONUP = 6,
                       EVENT_LBUTTONDBLCLK = 7, EVENT_RBUTTONDBLCLK = 8, EVENT_MBUTTONDBLCLK = 9, EVENT_MOUSEWHEEL = 10, EVENT_MOUSEHWHEEL = 11 };
enum ImreadModes { IMREAD_UNCHANGED = -1, IMREAD_GRAYSCALE = 0, IMREAD_COLOR = 1 };
} // namespace cv
