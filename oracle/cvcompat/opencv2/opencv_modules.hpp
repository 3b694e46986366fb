// ORACLE — TEST INFRASTRUCTURE ONLY. cvcompat: module list (none of the optional modules exist here).
#pragma once
#define HAVE_OPENCV_CORE
#define HAVE_OPENCV_IMGPROC
#define HAVE_OPENCV_HIGHGUI
#define HAVE_OPENCV_FEATURES2D
#define HAVE_OPENCV_VIDEO
