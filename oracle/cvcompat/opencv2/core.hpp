// ORACLE — TEST INFRASTRUCTURE ONLY.
// Minimal OpenCV-compatible surface ("cvcompat"): just enough of cv::Mat & friends for the UNMODIFIED reference sources of the
// hot path (modules/video/src/BackgroundSubtractor{SuBSENSE,LOBSTER,PAWCS,LBSP}.cpp, BackgroundSubtractionUtils.cpp,
// modules/features2d/src/LBSP.cpp and the litiv/utils headers they include) to compile where they lie under /root/reference, with
// g++ only. This is NOT OpenCV and not a copy of it: types are written from the public API documentation; only what the reference
// touches is implemented (2-D dense matrices, reference-counted), the rest is declared so that uninstantiated templates parse.
// The image-processing calls the reference makes inside apply() (imgproc.hpp) are implemented in cvcompat.cpp on top of the
// oracle's own mask operations, which tests/test_oracle_cpu.py pins bit-exactly against cv2 4.13.
// Built by oracle/Makefile target `_ref` into oracle/_ref/ (git-ignored); never part of the product.
#pragma once
#include <algorithm>
#include <cassert>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <initializer_list>
#include <iostream>
#include <memory>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

typedef unsigned char uchar;
typedef signed char schar;
typedef unsigned short ushort;
typedef int64_t int64;
typedef uint64_t uint64;

#define CV_VERSION_MAJOR 3
#define CV_VERSION_MINOR 4
#define CV_VERSION_REVISION 0
#define CV_MAJOR_VERSION CV_VERSION_MAJOR
#define CV_MINOR_VERSION CV_VERSION_MINOR

#define CV_CN_MAX 512
#define CV_CN_SHIFT 3
#define CV_DEPTH_MAX (1 << CV_CN_SHIFT)
#define CV_8U 0
#define CV_8S 1
#define CV_16U 2
#define CV_16S 3
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_MAT_DEPTH_MASK (CV_DEPTH_MAX - 1)
#define CV_MAT_DEPTH(flags) ((flags) & CV_MAT_DEPTH_MASK)
#define CV_MAKETYPE(depth, cn) (CV_MAT_DEPTH(depth) + (((cn) - 1) << CV_CN_SHIFT))
#define CV_MAKE_TYPE CV_MAKETYPE
#define CV_MAT_CN_MASK ((CV_CN_MAX - 1) << CV_CN_SHIFT)
#define CV_MAT_CN(flags) ((((flags) & CV_MAT_CN_MASK) >> CV_CN_SHIFT) + 1)
#define CV_MAT_TYPE_MASK (CV_DEPTH_MAX * CV_CN_MAX - 1)
#define CV_MAT_TYPE(flags) ((flags) & CV_MAT_TYPE_MASK)
#define CV_ELEM_SIZE1(type) ((int)(CV_MAT_DEPTH(type) <= 1 ? 1 : CV_MAT_DEPTH(type) < 4 ? 2 : CV_MAT_DEPTH(type) < 6 ? 4 : 8))
#define CV_ELEM_SIZE(type) (CV_MAT_CN(type) * CV_ELEM_SIZE1(type))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC2 CV_MAKETYPE(CV_8U, 2)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_8UC4 CV_MAKETYPE(CV_8U, 4)
#define CV_8UC(n) CV_MAKETYPE(CV_8U, (n))
#define CV_8SC1 CV_MAKETYPE(CV_8S, 1)
#define CV_8SC(n) CV_MAKETYPE(CV_8S, (n))
#define CV_16UC1 CV_MAKETYPE(CV_16U, 1)
#define CV_16UC2 CV_MAKETYPE(CV_16U, 2)
#define CV_16UC3 CV_MAKETYPE(CV_16U, 3)
#define CV_16UC4 CV_MAKETYPE(CV_16U, 4)
#define CV_16UC(n) CV_MAKETYPE(CV_16U, (n))
#define CV_16SC1 CV_MAKETYPE(CV_16S, 1)
#define CV_16SC(n) CV_MAKETYPE(CV_16S, (n))
#define CV_32SC1 CV_MAKETYPE(CV_32S, 1)
#define CV_32SC2 CV_MAKETYPE(CV_32S, 2)
#define CV_32SC(n) CV_MAKETYPE(CV_32S, (n))
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_32FC2 CV_MAKETYPE(CV_32F, 2)
#define CV_32FC3 CV_MAKETYPE(CV_32F, 3)
#define CV_32FC4 CV_MAKETYPE(CV_32F, 4)
#define CV_32FC(n) CV_MAKETYPE(CV_32F, (n))
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)
#define CV_64FC(n) CV_MAKETYPE(CV_64F, (n))
#define CV_Assert(expr) do { if(!(expr)) throw std::runtime_error(std::string("cvcompat: assertion failed: " #expr)); } while(0)
#define CV_DbgAssert(expr) ((void)0)
#define CV_Error(code, msg) throw std::runtime_error(std::string("cvcompat: ") + (msg))
#define CV_PI 3.1415926535897932384626433832795
#define CV_EXPORTS
#define CV_EXPORTS_W
#define CV_WRAP
#define CV_OUT
#define CV_IN_OUT

namespace cv {

typedef std::string String;

inline int cvRoundHalfEven(double v) { return (int)std::nearbyint(v); } // default rounding mode: to nearest, ties to even (== cvRound)
inline int cvRound(double v) { return cvRoundHalfEven(v); }
inline int cvFloor(double v) { return (int)std::floor(v); }
inline int cvCeil(double v) { return (int)std::ceil(v); }
inline size_t alignSize(size_t sz, int n) { return (sz + n - 1) & -n; }
int borderInterpolate(int p, int len, int borderType);

// ---- saturate_cast (documented semantics: round half to even for floats, then clamp)
template<typename T> inline T saturate_cast(uchar v) { return T(v); }
template<typename T> inline T saturate_cast(schar v) { return T(v); }
template<typename T> inline T saturate_cast(ushort v) { return T(v); }
template<typename T> inline T saturate_cast(short v) { return T(v); }
template<typename T> inline T saturate_cast(unsigned v) { return T(v); }
template<typename T> inline T saturate_cast(int v) { return T(v); }
template<typename T> inline T saturate_cast(float v) { return T(v); }
template<typename T> inline T saturate_cast(double v) { return T(v); }
template<typename T> inline T saturate_cast(int64 v) { return T(v); }
template<typename T> inline T saturate_cast(uint64 v) { return T(v); }
template<> inline uchar saturate_cast<uchar>(schar v) { return (uchar)std::max((int)v, 0); }
template<> inline uchar saturate_cast<uchar>(ushort v) { return (uchar)std::min((unsigned)v, (unsigned)UCHAR_MAX); }
template<> inline uchar saturate_cast<uchar>(int v) { return (uchar)((unsigned)v <= UCHAR_MAX ? v : v > 0 ? UCHAR_MAX : 0); }
template<> inline uchar saturate_cast<uchar>(short v) { return saturate_cast<uchar>((int)v); }
template<> inline uchar saturate_cast<uchar>(unsigned v) { return (uchar)std::min(v, (unsigned)UCHAR_MAX); }
template<> inline uchar saturate_cast<uchar>(float v) { return saturate_cast<uchar>(cvRound(v)); }
template<> inline uchar saturate_cast<uchar>(double v) { return saturate_cast<uchar>(cvRound(v)); }
template<> inline uchar saturate_cast<uchar>(int64 v) { return (uchar)((uint64)v <= (uint64)UCHAR_MAX ? v : v > 0 ? UCHAR_MAX : 0); }
template<> inline uchar saturate_cast<uchar>(uint64 v) { return (uchar)std::min(v, (uint64)UCHAR_MAX); }
template<> inline ushort saturate_cast<ushort>(int v) { return (ushort)((unsigned)v <= (unsigned)USHRT_MAX ? v : v > 0 ? USHRT_MAX : 0); }
template<> inline ushort saturate_cast<ushort>(float v) { return saturate_cast<ushort>(cvRound(v)); }
template<> inline ushort saturate_cast<ushort>(double v) { return saturate_cast<ushort>(cvRound(v)); }
template<> inline short saturate_cast<short>(int v) { return (short)((unsigned)(v - SHRT_MIN) <= (unsigned)USHRT_MAX ? v : v > 0 ? SHRT_MAX : SHRT_MIN); }
template<> inline short saturate_cast<short>(float v) { return saturate_cast<short>(cvRound(v)); }
template<> inline short saturate_cast<short>(double v) { return saturate_cast<short>(cvRound(v)); }
template<> inline int saturate_cast<int>(float v) { return cvRound(v); }
template<> inline int saturate_cast<int>(double v) { return cvRound(v); }
template<> inline schar saturate_cast<schar>(int v) { return (schar)((unsigned)(v - SCHAR_MIN) <= (unsigned)UCHAR_MAX ? v : v > 0 ? SCHAR_MAX : SCHAR_MIN); }
template<> inline schar saturate_cast<schar>(float v) { return saturate_cast<schar>(cvRound(v)); }
template<> inline schar saturate_cast<schar>(double v) { return saturate_cast<schar>(cvRound(v)); }

// ---- small fixed-size types
template<typename T, int m, int n> class Matx {
public:
    enum { rows = m, cols = n, channels = m * n };
    typedef T value_type;
    T val[m * n];
    Matx() { for(int i = 0; i < m * n; ++i) val[i] = T(0); }
    Matx(T v0) : Matx() { val[0] = v0; }
    Matx(T v0, T v1) : Matx() { static_assert(m * n >= 2, ""); val[0] = v0; val[1] = v1; }
    Matx(T v0, T v1, T v2) : Matx() { static_assert(m * n >= 3, ""); val[0] = v0; val[1] = v1; val[2] = v2; }
    Matx(T v0, T v1, T v2, T v3) : Matx() { static_assert(m * n >= 4, ""); val[0] = v0; val[1] = v1; val[2] = v2; val[3] = v3; }
    static Matx all(T a) { Matx r; for(int i = 0; i < m * n; ++i) r.val[i] = a; return r; }
    const T& operator()(int i, int j) const { return val[i * n + j]; }
    T& operator()(int i, int j) { return val[i * n + j]; }
    const T& operator()(int i) const { return val[i]; }
    T& operator()(int i) { return val[i]; }
    bool operator==(const Matx& o) const { for(int i = 0; i < m * n; ++i) if(!(val[i] == o.val[i])) return false; return true; }
    bool operator!=(const Matx& o) const { return !(*this == o); }
};
template<typename T, int cn> class Vec : public Matx<T, cn, 1> {
public:
    typedef T value_type;
    enum { channels = cn };
    using Matx<T, cn, 1>::Matx;
    Vec() {}
    Vec(const Matx<T, cn, 1>& a) : Matx<T, cn, 1>(a) {}
    static Vec all(T a) { Vec r; for(int i = 0; i < cn; ++i) r.val[i] = a; return r; }
    const T& operator[](int i) const { return this->val[i]; }
    T& operator[](int i) { return this->val[i]; }
    template<typename T2> operator Vec<T2, cn>() const { Vec<T2, cn> r; for(int i = 0; i < cn; ++i) r.val[i] = saturate_cast<T2>(this->val[i]); return r; }
};
typedef Vec<uchar, 2> Vec2b; typedef Vec<uchar, 3> Vec3b; typedef Vec<uchar, 4> Vec4b;
typedef Vec<short, 2> Vec2s; typedef Vec<short, 3> Vec3s; typedef Vec<short, 4> Vec4s;
typedef Vec<ushort, 2> Vec2w; typedef Vec<ushort, 3> Vec3w; typedef Vec<ushort, 4> Vec4w;
typedef Vec<int, 2> Vec2i; typedef Vec<int, 3> Vec3i; typedef Vec<int, 4> Vec4i;
typedef Vec<float, 2> Vec2f; typedef Vec<float, 3> Vec3f; typedef Vec<float, 4> Vec4f;
typedef Vec<double, 2> Vec2d; typedef Vec<double, 3> Vec3d; typedef Vec<double, 4> Vec4d;
template<typename T, int cn> inline Vec<T, cn> operator+(const Vec<T, cn>& a, const Vec<T, cn>& b) { Vec<T, cn> r; for(int i = 0; i < cn; ++i) r[i] = saturate_cast<T>(a[i] + b[i]); return r; }
template<typename T, int cn> inline Vec<T, cn> operator-(const Vec<T, cn>& a, const Vec<T, cn>& b) { Vec<T, cn> r; for(int i = 0; i < cn; ++i) r[i] = saturate_cast<T>(a[i] - b[i]); return r; }
template<typename T, int cn> inline Vec<T, cn> operator*(const Vec<T, cn>& a, double s) { Vec<T, cn> r; for(int i = 0; i < cn; ++i) r[i] = saturate_cast<T>(a[i] * s); return r; }
template<typename T, int cn> inline Vec<T, cn> operator*(double s, const Vec<T, cn>& a) { return a * s; }
template<typename T, int cn> inline Vec<T, cn> operator/(const Vec<T, cn>& a, double s) { Vec<T, cn> r; for(int i = 0; i < cn; ++i) r[i] = saturate_cast<T>(a[i] / s); return r; }
template<typename T, int cn> inline std::ostream& operator<<(std::ostream& os, const Vec<T, cn>& v) { os << "["; for(int i = 0; i < cn; ++i) os << (i ? ", " : "") << +v[i]; return os << "]"; }

template<typename T> class Point_ {
public:
    typedef T value_type;
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T _x, T _y) : x(_x), y(_y) {}
    template<typename T2> operator Point_<T2>() const { return Point_<T2>(saturate_cast<T2>(x), saturate_cast<T2>(y)); }
    bool operator==(const Point_& o) const { return x == o.x && y == o.y; }
    bool operator!=(const Point_& o) const { return !(*this == o); }
    Point_ operator+(const Point_& o) const { return Point_(x + o.x, y + o.y); }
    Point_ operator-(const Point_& o) const { return Point_(x - o.x, y - o.y); }
};
typedef Point_<int> Point2i; typedef Point_<int64> Point2l; typedef Point_<float> Point2f; typedef Point_<double> Point2d; typedef Point2i Point;
template<typename T> inline std::ostream& operator<<(std::ostream& os, const Point_<T>& p) { return os << "[" << p.x << ", " << p.y << "]"; }
template<typename T> class Point3_ { public: T x, y, z; Point3_() : x(0), y(0), z(0) {} Point3_(T a, T b, T c) : x(a), y(b), z(c) {} };
typedef Point3_<int> Point3i; typedef Point3_<float> Point3f; typedef Point3_<double> Point3d;

template<typename T> class Size_ {
public:
    typedef T value_type;
    T width, height;
    Size_() : width(0), height(0) {}
    Size_(T w, T h) : width(w), height(h) {}
    T area() const { return width * height; }
    bool empty() const { return width <= 0 || height <= 0; }
    bool operator==(const Size_& o) const { return width == o.width && height == o.height; }
    bool operator!=(const Size_& o) const { return !(*this == o); }
    template<typename T2> operator Size_<T2>() const { return Size_<T2>(saturate_cast<T2>(width), saturate_cast<T2>(height)); }
};
typedef Size_<int> Size2i; typedef Size_<int64> Size2l; typedef Size_<float> Size2f; typedef Size_<double> Size2d; typedef Size2i Size;
template<typename T> inline std::ostream& operator<<(std::ostream& os, const Size_<T>& s) { return os << "[" << s.width << " x " << s.height << "]"; }

template<typename T> class Rect_ {
public:
    typedef T value_type;
    T x, y, width, height;
    Rect_() : x(0), y(0), width(0), height(0) {}
    Rect_(T _x, T _y, T w, T h) : x(_x), y(_y), width(w), height(h) {}
    Rect_(const Point_<T>& p, const Size_<T>& s) : x(p.x), y(p.y), width(s.width), height(s.height) {}
    Point_<T> tl() const { return Point_<T>(x, y); }
    Point_<T> br() const { return Point_<T>(x + width, y + height); }
    Size_<T> size() const { return Size_<T>(width, height); }
    T area() const { return width * height; }
    bool empty() const { return width <= 0 || height <= 0; }
    bool contains(const Point_<T>& p) const { return x <= p.x && p.x < x + width && y <= p.y && p.y < y + height; }
    bool operator==(const Rect_& o) const { return x == o.x && y == o.y && width == o.width && height == o.height; }
    bool operator!=(const Rect_& o) const { return !(*this == o); }
};
typedef Rect_<int> Rect2i; typedef Rect_<float> Rect2f; typedef Rect_<double> Rect2d; typedef Rect2i Rect;
template<typename T> inline Rect_<T> operator&(const Rect_<T>& a, const Rect_<T>& b) {
    T x1 = std::max(a.x, b.x), y1 = std::max(a.y, b.y), x2 = std::min(a.x + a.width, b.x + b.width), y2 = std::min(a.y + a.height, b.y + b.height);
    return (x2 <= x1 || y2 <= y1) ? Rect_<T>() : Rect_<T>(x1, y1, x2 - x1, y2 - y1);
}

template<typename T> class Scalar_ : public Vec<T, 4> {
public:
    Scalar_() {}
    Scalar_(T v0) { this->val[0] = v0; }
    Scalar_(T v0, T v1, T v2 = 0, T v3 = 0) { this->val[0] = v0; this->val[1] = v1; this->val[2] = v2; this->val[3] = v3; }
    template<typename T2, int cn> Scalar_(const Vec<T2, cn>& v) { for(int i = 0; i < cn && i < 4; ++i) this->val[i] = saturate_cast<T>(v[i]); }
    static Scalar_ all(T v0) { return Scalar_(v0, v0, v0, v0); }
    template<typename T2> operator Scalar_<T2>() const { return Scalar_<T2>(saturate_cast<T2>(this->val[0]), saturate_cast<T2>(this->val[1]), saturate_cast<T2>(this->val[2]), saturate_cast<T2>(this->val[3])); }
};
typedef Scalar_<double> Scalar;

class Range {
public:
    int start, end;
    Range() : start(0), end(0) {}
    Range(int s, int e) : start(s), end(e) {}
    int size() const { return end - start; }
    bool empty() const { return start == end; }
    static Range all() { return Range(INT_MIN, INT_MAX); }
    bool operator==(const Range& o) const { return start == o.start && end == o.end; }
    bool operator!=(const Range& o) const { return !(*this == o); }
};

class KeyPoint {
public:
    Point2f pt; float size, angle, response; int octave, class_id;
    KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(Point2f p, float s, float a = -1, float r = 0, int o = 0, int c = -1) : pt(p), size(s), angle(a), response(r), octave(o), class_id(c) {}
    KeyPoint(float x, float y, float s, float a = -1, float r = 0, int o = 0, int c = -1) : pt(x, y), size(s), angle(a), response(r), octave(o), class_id(c) {}
};

// ---- type traits
template<typename T> class DataType { public: typedef T value_type; typedef T channel_type; enum { generic_type = 1, depth = -1, channels = 1, fmt = 0, type = -1 }; };
#define CVCOMPAT_DATATYPE(T, D) template<> class DataType<T> { public: typedef T value_type; typedef T channel_type; typedef T work_type; typedef T vec_type; \
    enum { generic_type = 0, depth = D, channels = 1, fmt = 0, type = CV_MAKETYPE(D, 1) }; };
CVCOMPAT_DATATYPE(bool, CV_8U) CVCOMPAT_DATATYPE(uchar, CV_8U) CVCOMPAT_DATATYPE(schar, CV_8S) CVCOMPAT_DATATYPE(char, CV_8S)
CVCOMPAT_DATATYPE(ushort, CV_16U) CVCOMPAT_DATATYPE(short, CV_16S) CVCOMPAT_DATATYPE(int, CV_32S) CVCOMPAT_DATATYPE(float, CV_32F) CVCOMPAT_DATATYPE(double, CV_64F)
template<typename T, int cn> class DataType<Vec<T, cn>> { public: typedef Vec<T, cn> value_type; typedef T channel_type;
    enum { generic_type = 0, depth = DataType<T>::depth, channels = cn, fmt = 0, type = CV_MAKETYPE(DataType<T>::depth, cn) }; };
template<typename T, int m, int n> class DataType<Matx<T, m, n>> { public: typedef Matx<T, m, n> value_type; typedef T channel_type;
    enum { generic_type = 0, depth = DataType<T>::depth, channels = m * n, fmt = 0, type = CV_MAKETYPE(DataType<T>::depth, m * n) }; };
template<typename T> class DataType<Point_<T>> { public: typedef Point_<T> value_type; typedef T channel_type;
    enum { generic_type = 0, depth = DataType<T>::depth, channels = 2, fmt = 0, type = CV_MAKETYPE(DataType<T>::depth, 2) }; };
template<typename T> class DataType<Scalar_<T>> { public: typedef Scalar_<T> value_type; typedef T channel_type;
    enum { generic_type = 0, depth = DataType<T>::depth, channels = 4, fmt = 0, type = CV_MAKETYPE(DataType<T>::depth, 4) }; };
namespace traits {
template<typename T> struct Type { enum { value = DataType<T>::type }; };
template<typename T> struct Depth { enum { value = DataType<T>::depth }; };
}

enum BorderTypes { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1, BORDER_REFLECT = 2, BORDER_WRAP = 3, BORDER_REFLECT_101 = 4, BORDER_TRANSPARENT = 5,
                   BORDER_REFLECT101 = BORDER_REFLECT_101, BORDER_DEFAULT = BORDER_REFLECT_101, BORDER_ISOLATED = 16 };
enum NormTypes { NORM_INF = 1, NORM_L1 = 2, NORM_L2 = 4, NORM_L2SQR = 5, NORM_HAMMING = 6, NORM_HAMMING2 = 7, NORM_TYPE_MASK = 7, NORM_RELATIVE = 8, NORM_MINMAX = 32 };
enum CmpTypes { CMP_EQ = 0, CMP_GT = 1, CMP_GE = 2, CMP_LT = 3, CMP_LE = 4, CMP_NE = 5 };
enum UMatUsageFlags { USAGE_DEFAULT = 0, USAGE_ALLOCATE_HOST_MEMORY = 1, USAGE_ALLOCATE_DEVICE_MEMORY = 2, USAGE_ALLOCATE_SHARED_MEMORY = 4 };
enum AccessFlag { ACCESS_READ = 1 << 24, ACCESS_WRITE = 1 << 25, ACCESS_RW = 3 << 24 };

// ---- allocator hooks (the reference's utils declare an aligned allocator; the compat Mat honours MatAllocator::allocate/deallocate)
class MatAllocator;
struct UMatData {
    enum { USER_ALLOCATED = 32 };
    explicit UMatData(const MatAllocator* a) : prevAllocator(a), currAllocator(a), urefcount(0), refcount(0), data(nullptr), origdata(nullptr), size(0), flags(0),
                                               handle(nullptr), userdata(nullptr), allocatorFlags_(0), mapcount(0), originalUMatData(nullptr) {}
    const MatAllocator* prevAllocator; const MatAllocator* currAllocator;
    int urefcount, refcount;
    uchar* data; uchar* origdata; size_t size;
    int flags;
    void* handle; void* userdata; int allocatorFlags_, mapcount; UMatData* originalUMatData;
};
class MatAllocator {
public:
    MatAllocator() {}
    virtual ~MatAllocator() {}
    virtual UMatData* allocate(int dims, const int* sizes, int type, void* data, size_t* step, int flags, UMatUsageFlags usageFlags) const = 0;
    virtual bool allocate(UMatData* data, int accessflags, UMatUsageFlags usageFlags) const = 0;
    virtual void deallocate(UMatData* data) const = 0;
};

struct MatSize {
    explicit MatSize(int* _p) : p(_p) {}
    int dims() const { return p[-1]; }
    Size operator()() const { return Size(p[1], p[0]); }
    const int& operator[](int i) const { return p[i]; }
    int& operator[](int i) { return p[i]; }
    operator const int*() const { return p; }
    bool operator==(const MatSize& o) const { int d = p[-1]; if(d != o.p[-1]) return false; if(d == 2) return p[0] == o.p[0] && p[1] == o.p[1]; for(int i = 0; i < d; ++i) if(p[i] != o.p[i]) return false; return true; }
    bool operator!=(const MatSize& o) const { return !(*this == o); }
    int* p;
};
struct MatStep {
    MatStep() { p = buf; buf[0] = buf[1] = 0; }
    explicit MatStep(size_t s) { p = buf; buf[0] = s; buf[1] = 0; }
    const size_t& operator[](int i) const { return p[i]; }
    size_t& operator[](int i) { return p[i]; }
    operator size_t() const { return buf[0]; }
    MatStep& operator=(size_t s) { buf[0] = s; return *this; }
    size_t* p; size_t buf[2];
private:
    MatStep(const MatStep&); MatStep& operator=(const MatStep&);
};

class _InputArray; class _OutputArray;
typedef const _InputArray& InputArray;
typedef InputArray InputArrayOfArrays;
typedef const _OutputArray& OutputArray;
typedef OutputArray OutputArrayOfArrays;
typedef OutputArray InputOutputArray;
typedef InputOutputArray InputOutputArrayOfArrays;
InputOutputArray noArray();
template<typename T> class Mat_;
class MatConstIterator; template<typename T> class MatConstIterator_; template<typename T> class MatIterator_;

/// 2-D dense matrix with shared, reference-counted storage (header copy = shallow copy, like cv::Mat)
class Mat {
public:
    enum { MAGIC_VAL = 0x42FF0000, AUTO_STEP = 0, CONTINUOUS_FLAG = 1 << 14, SUBMATRIX_FLAG = 1 << 15, MAGIC_MASK = 0xFFFF0000, TYPE_MASK = 0x00000FFF, DEPTH_MASK = 7 };
    Mat() : size(&rows) { reset_(); }
    Mat(int r, int c, int type) : size(&rows) { reset_(); create(r, c, type); }
    Mat(Size s, int type) : size(&rows) { reset_(); create(s.height, s.width, type); }
    Mat(int r, int c, int type, const Scalar& v) : size(&rows) { reset_(); create(r, c, type); *this = v; }
    Mat(Size s, int type, const Scalar& v) : size(&rows) { reset_(); create(s.height, s.width, type); *this = v; }
    Mat(int ndims, const int* sizes, int type) : size(&rows) { reset_(); create(ndims, sizes, type); }
    Mat(int r, int c, int type, void* d, size_t st = AUTO_STEP) : size(&rows) { reset_(); wrap_(r, c, type, d, st); }
    Mat(Size s, int type, void* d, size_t st = AUTO_STEP) : size(&rows) { reset_(); wrap_(s.height, s.width, type, d, st); }
    Mat(int ndims, const int* sizes, int type, void* d, const size_t* steps = nullptr) : size(&rows) {
        reset_(); CV_Assert(ndims >= 1 && ndims <= 2);
        if(ndims == 1) wrap_(sizes[0], 1, type, d, steps ? steps[0] : AUTO_STEP); else wrap_(sizes[0], sizes[1], type, d, steps ? steps[0] : AUTO_STEP);
    }
    Mat(const Mat& m) : size(&rows) { reset_(); assign_(m); }
    Mat(const Mat& m, const Rect& roi) : size(&rows) { reset_(); assign_(m); roi_(roi); }
    Mat(Mat&& m) : size(&rows) { reset_(); assign_(m); m.release(); }
    template<typename T> explicit Mat(const std::vector<T>& vec, bool copyData = false) : size(&rows) {
        reset_();
        if(vec.empty()) return;
        if(copyData) { create((int)vec.size(), 1, DataType<T>::type); std::memcpy(data, vec.data(), vec.size() * sizeof(T)); }
        else wrap_((int)vec.size(), 1, DataType<T>::type, (void*)vec.data(), AUTO_STEP);
    }
    ~Mat() { release(); }
    Mat& operator=(const Mat& m) { if(this != &m) { release(); assign_(m); } return *this; }
    Mat& operator=(Mat&& m) { if(this != &m) { release(); assign_(m); m.release(); } return *this; }
    Mat& operator=(const Scalar& s);
    Mat& setTo(InputArray value, InputArray mask);
    Mat& setTo(InputArray value);

    void create(int r, int c, int type);
    void create(Size s, int type) { create(s.height, s.width, type); }
    void create(int ndims, const int* sizes, int type) { CV_Assert(ndims >= 1 && ndims <= 2); create(sizes[0], ndims == 2 ? sizes[1] : 1, type); }
    void release() { owner_.reset(); udata_.reset(); reset_(); }
    Mat clone() const { Mat m; copyTo(m); return m; }
    void copyTo(Mat& m) const;
    void copyTo(OutputArray m) const;
    void copyTo(OutputArray m, InputArray mask) const;
    void convertTo(OutputArray m, int rtype, double alpha = 1, double beta = 0) const;
    Mat reshape(int cn, int rows_ = 0) const;
    Mat operator()(const Rect& roi) const { return Mat(*this, roi); }
    Mat operator()(const std::vector<Range>& ranges) const;
    Mat operator()(const Range* ranges) const;
    Mat operator()(Range rowRange, Range colRange) const;
    Mat mul(const Mat& m, double scale = 1) const;   // element-wise product (8-bit, saturating)
    Mat row(int y) const { return Mat(*this, Rect(0, y, cols, 1)); }
    Mat col(int x) const { return Mat(*this, Rect(x, 0, 1, rows)); }
    Mat rowRange(int a, int b) const { return Mat(*this, Rect(0, a, cols, b - a)); }
    Mat colRange(int a, int b) const { return Mat(*this, Rect(a, 0, b - a, rows)); }
    static Mat zeros(int r, int c, int type) { Mat m(r, c, type); if(m.data) for(int y = 0; y < r; ++y) std::memset(m.ptr(y), 0, (size_t)c * m.elemSize()); return m; }
    static Mat zeros(Size s, int type) { return zeros(s.height, s.width, type); }
    static Mat ones(int r, int c, int type) { return Mat(r, c, type, Scalar(1)); }
    static Mat ones(Size s, int type) { return Mat(s, type, Scalar(1)); }

    bool isContinuous() const { return (flags & CONTINUOUS_FLAG) != 0; }
    bool isSubmatrix() const { return (flags & SUBMATRIX_FLAG) != 0; }
    size_t elemSize() const { return (size_t)CV_ELEM_SIZE(flags); }
    size_t elemSize1() const { return (size_t)CV_ELEM_SIZE1(flags); }
    int type() const { return CV_MAT_TYPE(flags); }
    int depth() const { return CV_MAT_DEPTH(flags); }
    int channels() const { return CV_MAT_CN(flags); }
    size_t step1(int i = 0) const { return step.p[i] / elemSize1(); }
    bool empty() const { return data == nullptr || total() == 0; }
    size_t total() const { return (size_t)rows * cols; }
    int checkVector(int elemChannels, int depth_ = -1, bool requireContinuous = true) const;

    uchar* ptr(int y = 0) { return data + step.buf[0] * y; }
    const uchar* ptr(int y = 0) const { return data + step.buf[0] * y; }
    uchar* ptr(int y, int x) { return data + step.buf[0] * y + step.buf[1] * x; }
    const uchar* ptr(int y, int x) const { return data + step.buf[0] * y + step.buf[1] * x; }
    template<typename T> T* ptr(int y = 0) { return (T*)(data + step.buf[0] * y); }
    template<typename T> const T* ptr(int y = 0) const { return (const T*)(data + step.buf[0] * y); }
    template<typename T> T* ptr(int y, int x) { return (T*)(data + step.buf[0] * y + step.buf[1] * x); }
    template<typename T> const T* ptr(int y, int x) const { return (const T*)(data + step.buf[0] * y + step.buf[1] * x); }
    template<typename T> T& at(int y, int x) { return ((T*)(data + step.buf[0] * y))[x]; }
    template<typename T> const T& at(int y, int x) const { return ((const T*)(data + step.buf[0] * y))[x]; }
    template<typename T> T& at(int i) { if(isContinuous() || rows == 1) return ((T*)data)[i]; if(cols == 1) return *(T*)(data + step.buf[0] * i); return at<T>(i / cols, i % cols); }
    template<typename T> const T& at(int i) const { return const_cast<Mat*>(this)->at<T>(i); }
    template<typename T> T& at(Point p) { return at<T>(p.y, p.x); }
    template<typename T> const T& at(Point p) const { return at<T>(p.y, p.x); }
    template<typename T> T& at(const int* idx) { return at<T>(idx[0], idx[1]); }
    template<typename T> const T& at(const int* idx) const { return at<T>(idx[0], idx[1]); }
    template<typename T> MatIterator_<T> begin();
    template<typename T> MatIterator_<T> end();
    template<typename T> MatConstIterator_<T> begin() const;
    template<typename T> MatConstIterator_<T> end() const;

    int flags;
    int dims;          // kept right before rows so that size.p[-1] == dims (cv::Mat layout contract used by cv::MatSize)
    int rows, cols;
    uchar* data;
    const uchar* datastart; const uchar* dataend; const uchar* datalimit;
    MatAllocator* allocator;
    UMatData* u;
    MatSize size;
    MatStep step;

protected:
    std::shared_ptr<uchar> owner_;     // storage allocated by create()
    std::shared_ptr<UMatData> udata_;  // storage obtained from a user MatAllocator
    void reset_() { flags = MAGIC_VAL; dims = 0; rows = cols = 0; data = nullptr; datastart = dataend = datalimit = nullptr; allocator = nullptr; u = nullptr; step.buf[0] = step.buf[1] = 0; }
    void assign_(const Mat& m) {
        flags = m.flags; dims = m.dims; rows = m.rows; cols = m.cols; data = m.data; datastart = m.datastart; dataend = m.dataend; datalimit = m.datalimit;
        allocator = m.allocator; u = m.u; step.buf[0] = m.step.buf[0]; step.buf[1] = m.step.buf[1]; owner_ = m.owner_; udata_ = m.udata_;
    }
    void wrap_(int r, int c, int type, void* d, size_t st) {
        flags = MAGIC_VAL | CV_MAT_TYPE(type); dims = 2; rows = r; cols = c; data = (uchar*)d;
        const size_t esz = (size_t)CV_ELEM_SIZE(type), minstep = esz * c;
        step.buf[0] = (st == AUTO_STEP) ? minstep : st; step.buf[1] = esz;
        if(step.buf[0] == minstep || r <= 1) flags |= CONTINUOUS_FLAG;
        datastart = data; dataend = datalimit = data + step.buf[0] * (size_t)r;
    }
    void roi_(const Rect& roi) {
        CV_Assert(roi.x >= 0 && roi.y >= 0 && roi.width >= 0 && roi.height >= 0 && roi.x + roi.width <= cols && roi.y + roi.height <= rows);
        data += step.buf[0] * roi.y + step.buf[1] * roi.x;
        if(roi.width < cols || roi.height < rows) flags |= SUBMATRIX_FLAG;
        if(roi.width < cols && roi.height > 1) flags &= ~CONTINUOUS_FLAG;
        rows = roi.height; cols = roi.width;
        if(rows <= 1) flags |= CONTINUOUS_FLAG;
    }
};
static_assert(offsetof(Mat, rows) == offsetof(Mat, dims) + sizeof(int), "MatSize relies on dims sitting right before rows");

template<typename T> class Mat_ : public Mat {
public:
    typedef T value_type;
    typedef typename DataType<T>::channel_type channel_type;
    Mat_() : Mat() { flags = (flags & ~CV_MAT_TYPE_MASK) | DataType<T>::type; }
    Mat_(int r, int c) : Mat(r, c, DataType<T>::type) {}
    Mat_(int r, int c, const T& v) : Mat(r, c, DataType<T>::type) { *this = v; }
    explicit Mat_(Size s) : Mat(s, DataType<T>::type) {}
    Mat_(Size s, const T& v) : Mat(s, DataType<T>::type) { *this = v; }
    Mat_(int ndims, const int* sizes) : Mat(ndims, sizes, DataType<T>::type) {}
    Mat_(int r, int c, T* d, size_t st = AUTO_STEP) : Mat(r, c, DataType<T>::type, d, st) {}
    Mat_(int ndims, const int* sizes, T* d, const size_t* steps = nullptr) : Mat(ndims, sizes, DataType<T>::type, d, steps) {}
    Mat_(const Mat& m) : Mat() { flags = (flags & ~CV_MAT_TYPE_MASK) | DataType<T>::type; *this = m; }
    Mat_(const Mat_& m) : Mat(m) {}
    Mat_(const Mat_& m, const Rect& roi) : Mat(m, roi) {}
    Mat_& operator=(const Mat& m) {
        if(m.empty()) { release(); flags = (flags & ~CV_MAT_TYPE_MASK) | DataType<T>::type; return *this; }
        if(DataType<T>::type == m.type()) { Mat::operator=(m); return *this; }
        if(DataType<T>::depth == m.depth()) { Mat::operator=(m.reshape(DataType<T>::channels)); return *this; }
        Mat tmp; m.convertTo(tmp, DataType<T>::type); Mat::operator=(tmp); return *this;
    }
    Mat_& operator=(const Mat_& m) { Mat::operator=(m); return *this; }
    Mat_& operator=(const T& v) { for(int y = 0; y < rows; ++y) { T* p = (T*)ptr(y); for(int x = 0; x < cols; ++x) p[x] = v; } return *this; }
    void create(int r, int c) { Mat::create(r, c, DataType<T>::type); }
    void create(Size s) { Mat::create(s, DataType<T>::type); }
    void create(int ndims, const int* sizes) { Mat::create(ndims, sizes, DataType<T>::type); }
    Mat_ clone() const { return Mat_(Mat::clone()); }
    Mat_ operator()(const Rect& roi) const { return Mat_(*this, roi); }
    Mat_ operator()(const Range& rowRange, const Range& colRange) const { return Mat_(Mat::operator()(rowRange, colRange)); }
    static Mat_ zeros(Size s) { return Mat_(Mat::zeros(s, DataType<T>::type)); }
    static Mat_ ones(Size s) { Mat_ m(s); m = T(1); return m; }
    Mat_ operator()(const std::vector<Range>& ranges) const { return Mat_(Mat::operator()(ranges)); }
    int type() const { return DataType<T>::type; }
    int depth() const { return DataType<T>::depth; }
    int channels() const { return DataType<T>::channels; }
    size_t elemSize() const { return sizeof(T); }
    T* operator[](int y) { return (T*)ptr(y); }
    const T* operator[](int y) const { return (const T*)ptr(y); }
    T& operator()(int y, int x) { return ((T*)ptr(y))[x]; }
    const T& operator()(int y, int x) const { return ((const T*)ptr(y))[x]; }
    T& operator()(int i) { return this->template at<T>(i); }
    const T& operator()(int i) const { return this->template at<T>(i); }
    T& operator()(Point p) { return (*this)(p.y, p.x); }
    const T& operator()(Point p) const { return (*this)(p.y, p.x); }
    T& operator()(const int* idx) { return (*this)(idx[0], idx[1]); }
    const T& operator()(const int* idx) const { return (*this)(idx[0], idx[1]); }
    MatIterator_<T> begin(); MatIterator_<T> end();
    MatConstIterator_<T> begin() const; MatConstIterator_<T> end() const;
};
typedef Mat_<uchar> Mat1b; typedef Mat_<Vec3b> Mat3b; typedef Mat_<float> Mat1f; typedef Mat_<int> Mat1i; typedef Mat_<double> Mat1d;

// iterators / sparse matrices: only named by uninstantiated templates of the reference's utility headers
class MatConstIterator { public: const Mat* m; size_t elemSize; const uchar* ptr; MatConstIterator() : m(nullptr), elemSize(0), ptr(nullptr) {} void pos(int* idx) const; Point pos() const; };
template<typename T> class MatConstIterator_ : public MatConstIterator { public: const T& operator*() const { return *(const T*)ptr; } MatConstIterator_& operator++(); bool operator!=(const MatConstIterator_& o) const { return ptr != o.ptr; } bool operator==(const MatConstIterator_& o) const { return ptr == o.ptr; } };
template<typename T> class MatIterator_ : public MatConstIterator_<T> { public: T& operator*() const { return *(T*)this->ptr; } MatIterator_& operator++(); };
class SparseMat {
public:
    struct Node { size_t hashval, next; int idx[32]; };
    SparseMat(); SparseMat(int dims, const int* sizes, int type);
    void create(int dims, const int* sizes, int type); void clear(); size_t nzcount() const; int type() const; int dims() const; const int* size() const;
};
template<typename T> class SparseMat_ : public SparseMat { public: SparseMat_(); T& ref(const int* idx, size_t* hashval = nullptr); };
class SparseMatConstIterator { public: const SparseMat::Node* node() const; };

// ---- argument proxies
class _InputArray {
public:
    enum KindFlag { KIND_SHIFT = 16, FIXED_TYPE = 0x8000 << KIND_SHIFT, FIXED_SIZE = 0x4000 << KIND_SHIFT, KIND_MASK = 31 << KIND_SHIFT,
                    NONE = 0 << KIND_SHIFT, MAT = 1 << KIND_SHIFT, MATX = 2 << KIND_SHIFT, STD_VECTOR = 3 << KIND_SHIFT, STD_VECTOR_VECTOR = 4 << KIND_SHIFT,
                    STD_VECTOR_MAT = 5 << KIND_SHIFT, EXPR = 6 << KIND_SHIFT, OPENGL_BUFFER = 7 << KIND_SHIFT, CUDA_HOST_MEM = 8 << KIND_SHIFT,
                    CUDA_GPU_MAT = 9 << KIND_SHIFT, UMAT = 10 << KIND_SHIFT, STD_VECTOR_UMAT = 11 << KIND_SHIFT };
    _InputArray() : kind_(NONE), obj(nullptr), fixedType_(false) {}
    _InputArray(const Mat& m) : kind_(MAT), obj((void*)&m), fixedType_(false) {}
    template<typename T> _InputArray(const Mat_<T>& m) : kind_(MAT), obj((void*)static_cast<const Mat*>(&m)), fixedType_(true) {}
    _InputArray(const std::vector<Mat>& v) : kind_(STD_VECTOR_MAT), obj((void*)&v), fixedType_(false) {}
    template<typename T> _InputArray(const std::vector<Mat_<T>>& v);
    _InputArray(const double& v) : kind_(MATX), obj(nullptr), fixedType_(true), scalar_(v) {}
    template<typename T> _InputArray(const Scalar_<T>& s) : kind_(MATX), obj(nullptr), fixedType_(true), scalar_(s) {}
    Mat getMat(int i = -1) const;
    void getMatVector(std::vector<Mat>& mv) const;
    int kind() const { return kind_; }
    bool isMat() const { return kind_ == MAT; }
    bool isMatVector() const { return kind_ == STD_VECTOR_MAT; }
    bool isScalar_() const { return kind_ == MATX; }
    const Scalar& scalar_value_() const { return scalar_; }
    bool empty() const;
    Size size(int i = -1) const { return getMat(i).size(); }
    int type(int i = -1) const { return getMat(i).type(); }
    int depth(int i = -1) const { return getMat(i).depth(); }
    int channels(int i = -1) const { return getMat(i).channels(); }
    size_t total(int i = -1) const { return getMat(i).total(); }
    int rows(int i = -1) const { return getMat(i).rows; }
    int cols(int i = -1) const { return getMat(i).cols; }
    int dims(int i = -1) const { return getMat(i).dims; }
    bool isContinuous(int i = -1) const { return getMat(i).isContinuous(); }
    bool fixedType() const { return fixedType_; }
    void* getObj() const { return obj; }
protected:
    int kind_; void* obj; bool fixedType_; Scalar scalar_;
};
class _OutputArray : public _InputArray {
public:
    _OutputArray() {}
    _OutputArray(Mat& m) : _InputArray(m) {}
    template<typename T> _OutputArray(Mat_<T>& m) : _InputArray(m) {}
    _OutputArray(std::vector<Mat>& v) : _InputArray(v) {}
    template<typename T> _OutputArray(std::vector<Mat_<T>>& v);
    _OutputArray(const Mat& m) : _InputArray(m) {}   // cv allows writing through a const header (fixed size)
    bool needed() const { return kind_ != NONE; }
    Mat& getMatRef(int i = -1) const;
    void create(Size sz, int type, int i = -1, bool allowTransposed = false, int fixedDepthMask = 0) const;
    void create(int rows, int cols, int type, int i = -1, bool allowTransposed = false, int fixedDepthMask = 0) const;
    void create(int dims, const int* size, int type, int i = -1, bool allowTransposed = false, int fixedDepthMask = 0) const;
    void release() const;
    void setTo(const _InputArray& value, const _InputArray& mask = _InputArray()) const;
    void assign(const Mat& m) const;
};

// element-wise comparison against a scalar / another matrix: 8UC1 result with 255 where true (the reference uses `m==255`)
Mat operator==(const Mat& a, double s);
Mat operator!=(const Mat& a, double s);
Mat operator>(const Mat& a, double s);
Mat operator<(const Mat& a, double s);
Mat operator>=(const Mat& a, double s);
Mat operator<=(const Mat& a, double s);
Mat operator==(const Mat& a, const Mat& b);
Mat operator!=(const Mat& a, const Mat& b);
// the few matrix expressions the reference writes (cv evaluates them lazily through MatExpr; here they are evaluated at once)
Mat operator&(const Mat& a, const Mat& b);
Mat operator|(const Mat& a, const Mat& b);
Mat operator^(const Mat& a, const Mat& b);
Mat operator~(const Mat& a);
Mat operator/(const Mat& a, double s);   // == a.convertTo(., a.type(), 1/s): float multiply, round half to even, saturate
Mat operator*(const Mat& a, double s);
inline Mat operator*(double s, const Mat& a) { return a * s; }
Mat& operator+=(Mat& a, const Mat& b);   // saturating element-wise add (same type)
Mat& operator|=(Mat& a, const Scalar& s);
Mat& operator&=(Mat& a, const Scalar& s);
inline Mat& operator|=(Mat& a, int s) { return a |= Scalar::all((double)s); }
inline Mat& operator&=(Mat& a, int s) { return a &= Scalar::all((double)s); }

// ---- persistence / algorithm base (never exercised on the hot path; declared for the class hierarchy)
class FileNode { public: FileNode() {} bool empty() const { return true; } FileNode operator[](const char*) const { return FileNode(); } FileNode operator[](const String&) const { return FileNode(); }
    template<typename T> void operator>>(T&) const {} };
class FileStorage {
public:
    enum Mode { READ = 0, WRITE = 1, APPEND = 2, MEMORY = 4 };
    FileStorage() {} FileStorage(const String&, int, const String& = String()) {}
    virtual ~FileStorage() {}
    virtual bool open(const String&, int, const String& = String()) { return false; }
    virtual bool isOpened() const { return false; }
    virtual void release() {}
    FileNode operator[](const char*) const { return FileNode(); }
    FileNode operator[](const String&) const { return FileNode(); }
    FileNode root(int = 0) const { return FileNode(); }
};
template<typename T> inline FileStorage& operator<<(FileStorage& fs, const T&) { return fs; }
template<typename T> struct Ptr : public std::shared_ptr<T> { using std::shared_ptr<T>::shared_ptr; Ptr() {} Ptr(const std::shared_ptr<T>& o) : std::shared_ptr<T>(o) {} bool empty() const { return !*this; } };
template<typename T, typename... A> inline Ptr<T> makePtr(A&&... a) { return Ptr<T>(std::make_shared<T>(std::forward<A>(a)...)); }
class Algorithm {
public:
    Algorithm() {} virtual ~Algorithm() {}
    virtual void clear() {}
    virtual void write(FileStorage&) const {}
    virtual void read(const FileNode&) {}
    virtual bool empty() const { return false; }
    virtual void save(const String&) const {}
    virtual String getDefaultName() const { return "my_object"; }
};

// ---- core array operations used by the reference (cvcompat.cpp)
void bitwise_and(InputArray a, InputArray b, OutputArray dst, InputArray mask = noArray());
void bitwise_or(InputArray a, InputArray b, OutputArray dst, InputArray mask = noArray());
void bitwise_xor(InputArray a, InputArray b, OutputArray dst, InputArray mask = noArray());
void bitwise_not(InputArray a, OutputArray dst, InputArray mask = noArray());
int countNonZero(InputArray a);
Scalar sum(InputArray a);
Scalar mean(InputArray a, InputArray mask = noArray());
void split(const Mat& src, Mat* mv);
void split(InputArray src, OutputArrayOfArrays mv);
void merge(const Mat* mv, size_t count, OutputArray dst);
void merge(InputArrayOfArrays mv, OutputArray dst);
void max(InputArray a, InputArray b, OutputArray dst);
void min(InputArray a, InputArray b, OutputArray dst);
void max(const Mat& a, const Mat& b, Mat& dst);
void min(const Mat& a, const Mat& b, Mat& dst);
void absdiff(InputArray a, InputArray b, OutputArray dst);
void compare(InputArray a, InputArray b, OutputArray dst, int cmpop);
void normalize(InputArray src, InputOutputArray dst, double alpha = 1, double beta = 0, int norm_type = NORM_L2, int dtype = -1, InputArray mask = noArray());
void addWeighted(InputArray a, double alpha, InputArray b, double beta, double gamma, OutputArray dst, int dtype = -1);
// only named by templates of the reference's umbrella headers that the compiled sources never instantiate (litiv/imgproc.hpp)
struct TermCriteria { enum { COUNT = 1, MAX_ITER = 1, EPS = 2 }; int type, maxCount; double epsilon; TermCriteria(int t = 0, int n = 0, double e = 0) : type(t), maxCount(n), epsilon(e) {} };
#define CV_TERMCRIT_ITER 1
#define CV_TERMCRIT_EPS 2
enum KmeansFlags { KMEANS_RANDOM_CENTERS = 0, KMEANS_PP_CENTERS = 2, KMEANS_USE_INITIAL_LABELS = 1 };
double kmeans(InputArray data, int K, InputOutputArray bestLabels, TermCriteria criteria, int attempts, int flags, OutputArray centers = noArray());
void convertScaleAbs(InputArray src, OutputArray dst, double alpha = 1, double beta = 0);
void minMaxIdx(InputArray src, double* minVal, double* maxVal = nullptr, int* minIdx = nullptr, int* maxIdx = nullptr, InputArray mask = noArray());
void minMaxLoc(InputArray src, double* minVal, double* maxVal = nullptr, Point* minLoc = nullptr, Point* maxLoc = nullptr, InputArray mask = noArray());
double norm(InputArray a, int normType = NORM_L2, InputArray mask = noArray());
double norm(InputArray a, InputArray b, int normType = NORM_L2, InputArray mask = noArray());
double determinant(InputArray m);
double invert(InputArray src, OutputArray dst, int flags = 0);
void copyMakeBorder(InputArray src, OutputArray dst, int top, int bottom, int left, int right, int borderType, const Scalar& value = Scalar());
template<typename T, int m, int n> double determinant(const Matx<T, m, n>&);

using std::max; using std::min; using std::abs; using std::sqrt; using std::swap;

} // namespace cv
