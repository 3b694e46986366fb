// ORACLE — TEST INFRASTRUCTURE ONLY.
// Implementation of the cvcompat surface (oracle/cvcompat/opencv2/*.hpp): cv::Mat storage and the handful of array / image
// operations the reference's hot path calls. The mask operations (morphology, median, flood fill, INTER_AREA resize) forward to
// the oracle's own implementations in ../lvo_common.hpp, which tests/test_oracle_cpu.py pins bit-exactly against cv2 4.13;
// the float ops follow the cv2-probed semantics of SURVEY.md Appendix E. Anything the hot path does not need throws.
#include "opencv2/core.hpp"
#include "opencv2/imgproc.hpp"
#include "opencv2/highgui.hpp"
#include "opencv2/features2d.hpp"
#include "../lvo_common.hpp"
#include "../lvo_edge_lbsp.hpp"   // normalize_minmax_u8 (the restatement of cv::normalize pinned against cv2)

namespace cv {

static void unsupported(const char* what) { throw std::runtime_error(std::string("cvcompat: unsupported: ") + what); }

int borderInterpolate(int p, int len, int borderType) {
    if((unsigned)p < (unsigned)len) return p;
    if(borderType == BORDER_REPLICATE) return p < 0 ? 0 : len - 1;
    if(borderType == BORDER_REFLECT || borderType == BORDER_REFLECT_101) {
        const int delta = borderType == BORDER_REFLECT_101;
        if(len == 1) return 0;
        do { if(p < 0) p = -p - 1 + delta; else p = len - 1 - (p - len) - delta; } while((unsigned)p >= (unsigned)len);
        return p;
    }
    if(borderType == BORDER_WRAP) { if(p < 0) p -= ((p - len + 1) / len) * len; if(p >= len) p %= len; return p; }
    if(borderType == BORDER_CONSTANT) return -1;
    unsupported("borderInterpolate type");
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// Mat
// ---------------------------------------------------------------------------------------------------------------------
void Mat::create(int r, int c, int type) {
    type = CV_MAT_TYPE(type);
    if(data && rows == r && cols == c && this->type() == type && dims == 2) return;
    MatAllocator* a = allocator;
    release();
    allocator = a;
    CV_Assert(r >= 0 && c >= 0);
    flags = MAGIC_VAL | type | CONTINUOUS_FLAG; dims = 2; rows = r; cols = c;
    const size_t esz = (size_t)CV_ELEM_SIZE(type);
    step.buf[1] = esz; step.buf[0] = esz * c;
    const size_t bytes = step.buf[0] * (size_t)r;
    if(bytes == 0) return;
    if(allocator) {
        int sizes[2] = {r, c}; size_t steps[2] = {0, 0};
        UMatData* ud = allocator->allocate(2, sizes, type, nullptr, steps, 0, USAGE_DEFAULT);
        CV_Assert(ud && ud->data);
        const MatAllocator* al = allocator;
        udata_ = std::shared_ptr<UMatData>(ud, [al](UMatData* p) { al->deallocate(p); });
        u = ud; data = ud->data; step.buf[0] = steps[0]; step.buf[1] = steps[1];
        if(step.buf[0] != esz * c && r > 1) flags &= ~CONTINUOUS_FLAG;
    } else {
        void* p = nullptr;
        if(posix_memalign(&p, 64, bytes + 64) != 0) throw std::bad_alloc();
        owner_ = std::shared_ptr<uchar>((uchar*)p, [](uchar* q) { std::free(q); });
        data = (uchar*)p;
    }
    datastart = data; dataend = datalimit = data + step.buf[0] * (size_t)r;
}

void Mat::copyTo(Mat& m) const {
    if(empty()) { m.release(); return; }
    if(m.data == data && m.rows == rows && m.cols == cols && m.type() == type() && m.step.buf[0] == step.buf[0]) return;
    m.create(rows, cols, type());
    const size_t rowbytes = (size_t)cols * elemSize();
    for(int y = 0; y < rows; ++y) std::memmove(m.ptr(y), ptr(y), rowbytes);
}
void Mat::copyTo(OutputArray o) const {
    if(empty()) { o.release(); return; }
    o.create(rows, cols, type());
    Mat& m = o.getMatRef();
    copyTo(m);
}
void Mat::copyTo(OutputArray o, InputArray mask_) const {
    Mat mask = mask_.getMat();
    if(mask.empty()) { copyTo(o); return; }
    CV_Assert(mask.type() == CV_8UC1 && mask.rows == rows && mask.cols == cols);
    Mat& m = o.getMatRef();
    if(m.rows != rows || m.cols != cols || m.type() != type()) { m.create(rows, cols, type()); m = Scalar::all(0); }
    const size_t esz = elemSize();
    for(int y = 0; y < rows; ++y) {
        const uchar* s = ptr(y); uchar* d = m.ptr(y); const uchar* k = mask.ptr(y);
        for(int x = 0; x < cols; ++x) if(k[x]) std::memcpy(d + esz * x, s + esz * x, esz);
    }
}

template<typename T> static inline double load_as_double(const uchar* p) { return (double)*(const T*)p; }
static inline double get_elem(const uchar* p, int depth) {
    switch(depth) {
        case CV_8U: return load_as_double<uchar>(p); case CV_8S: return load_as_double<schar>(p); case CV_16U: return load_as_double<ushort>(p);
        case CV_16S: return load_as_double<short>(p); case CV_32S: return load_as_double<int>(p); case CV_32F: return load_as_double<float>(p);
        default: return load_as_double<double>(p);
    }
}
static inline void set_elem(uchar* p, int depth, double v) {
    switch(depth) {
        case CV_8U: *p = saturate_cast<uchar>(v); break; case CV_8S: *(schar*)p = saturate_cast<schar>(v); break;
        case CV_16U: *(ushort*)p = saturate_cast<ushort>(v); break; case CV_16S: *(short*)p = saturate_cast<short>(v); break;
        case CV_32S: *(int*)p = saturate_cast<int>(v); break; case CV_32F: *(float*)p = (float)v; break; default: *(double*)p = v;
    }
}

/// convertTo: without scaling an exact conversion (saturate_cast: round half to even, clamp); with scaling OpenCV computes
/// `src*alpha + beta` in float for 8/16-bit and float sources (double for 32S / 64F) before the saturating cast
void Mat::convertTo(OutputArray o, int rtype, double alpha, double beta) const {
    if(empty()) { o.release(); return; }
    const int sdepth = depth(), ddepth = rtype < 0 ? sdepth : CV_MAT_DEPTH(rtype), cn = channels();
    Mat src = *this; // keeps the storage alive when o aliases this
    Mat dst;
    if(o.getObj() == (void*)this && ddepth == sdepth) dst = src; else dst.create(rows, cols, CV_MAKETYPE(ddepth, cn));
    const bool noscale = alpha == 1.0 && beta == 0.0;
    const bool in_float = sdepth != CV_32S && sdepth != CV_64F && ddepth != CV_64F;
    const size_t s1 = src.elemSize1(), d1 = dst.elemSize1();
    for(int y = 0; y < rows; ++y) {
        const uchar* s = src.ptr(y); uchar* d = dst.ptr(y);
        for(int x = 0; x < cols * cn; ++x) {
            double v = get_elem(s + s1 * x, sdepth);
            if(!noscale) v = in_float ? (double)((float)v * (float)alpha + (float)beta) : v * alpha + beta;
            set_elem(d + d1 * x, ddepth, v);
        }
    }
    if(o.getObj() != (void*)this || ddepth != sdepth) o.getMatRef() = dst;
}

Mat Mat::reshape(int cn, int rows_) const {
    CV_Assert(rows_ == 0);
    if(cn == 0 || cn == channels()) return *this;
    const int total_ch = cols * channels();
    CV_Assert(total_ch % cn == 0 && (isContinuous() || rows == 1 || true));
    Mat m = *this;
    m.flags = (m.flags & ~CV_MAT_TYPE_MASK) | CV_MAKETYPE(depth(), cn);
    m.cols = total_ch / cn;
    m.step.buf[1] = m.elemSize();
    return m;
}
Mat Mat::operator()(const std::vector<Range>&) const { unsupported("Mat::operator()(ranges)"); return Mat(); }
Mat Mat::operator()(const Range*) const { unsupported("Mat::operator()(ranges)"); return Mat(); }
Mat Mat::operator()(Range r, Range c) const {
    const int y0 = r == Range::all() ? 0 : r.start, y1 = r == Range::all() ? rows : r.end, x0 = c == Range::all() ? 0 : c.start, x1 = c == Range::all() ? cols : c.end;
    return Mat(*this, Rect(x0, y0, x1 - x0, y1 - y0));
}
int Mat::checkVector(int, int, bool) const { unsupported("Mat::checkVector"); return -1; }

Mat& Mat::operator=(const Scalar& s) {
    const int cn = channels(), dp = depth(); const size_t e1 = elemSize1();
    for(int y = 0; y < rows; ++y) {
        uchar* d = ptr(y);
        for(int x = 0; x < cols; ++x) for(int c = 0; c < cn; ++c) set_elem(d + e1 * ((size_t)x * cn + c), dp, c < 4 ? s.val[c] : 0.0);
    }
    return *this;
}
Mat& Mat::setTo(InputArray value, InputArray mask_) {
    CV_Assert(value.isScalar_());
    const Scalar s = value.scalar_value_();
    Mat mask = mask_.getMat();
    if(mask.empty()) return *this = s;
    CV_Assert(mask.type() == CV_8UC1 && mask.rows == rows && mask.cols == cols);
    const int cn = channels(), dp = depth(); const size_t e1 = elemSize1();
    for(int y = 0; y < rows; ++y) {
        uchar* d = ptr(y); const uchar* k = mask.ptr(y);
        for(int x = 0; x < cols; ++x) if(k[x]) for(int c = 0; c < cn; ++c) set_elem(d + e1 * ((size_t)x * cn + c), dp, c < 4 ? s.val[c] : 0.0);
    }
    return *this;
}
Mat& Mat::setTo(InputArray value) { return setTo(value, noArray()); }

// ---------------------------------------------------------------------------------------------------------------------
// argument proxies
// ---------------------------------------------------------------------------------------------------------------------
static const _OutputArray g_none;
InputOutputArray noArray() { return g_none; }

Mat _InputArray::getMat(int i) const {
    if(kind_ == MAT) return *(const Mat*)obj;
    if(kind_ == NONE) return Mat();
    if(kind_ == STD_VECTOR_MAT) { const std::vector<Mat>& v = *(const std::vector<Mat>*)obj; CV_Assert(i >= 0 && (size_t)i < v.size()); return v[i]; }
    if(kind_ == MATX) { Mat m(4, 1, CV_64F); for(int k = 0; k < 4; ++k) m.at<double>(k) = scalar_.val[k]; return m; }
    unsupported("_InputArray kind"); return Mat();
}
void _InputArray::getMatVector(std::vector<Mat>& mv) const {
    if(kind_ == STD_VECTOR_MAT) { mv = *(const std::vector<Mat>*)obj; return; }
    if(kind_ == MAT) { mv.assign(1, *(const Mat*)obj); return; }
    mv.clear();
}
bool _InputArray::empty() const {
    if(kind_ == MAT) return ((const Mat*)obj)->empty();
    if(kind_ == STD_VECTOR_MAT) return ((const std::vector<Mat>*)obj)->empty();
    return kind_ == NONE;
}
Mat& _OutputArray::getMatRef(int i) const {
    if(kind_ == MAT) return *(Mat*)obj;
    if(kind_ == STD_VECTOR_MAT) { std::vector<Mat>& v = *(std::vector<Mat>*)obj; CV_Assert(i >= 0 && (size_t)i < v.size()); return v[i]; }
    unsupported("_OutputArray::getMatRef on a non-matrix"); static Mat dummy; return dummy;
}
void _OutputArray::create(int r, int c, int type, int i, bool, int) const {
    if(kind_ == NONE) return;
    if(kind_ == STD_VECTOR_MAT && i < 0) { ((std::vector<Mat>*)obj)->resize((size_t)r * c); return; }
    Mat& m = getMatRef(i);
    if(fixedType_) CV_Assert(CV_MAT_TYPE(type) == m.type() || m.empty());
    m.create(r, c, type);
}
void _OutputArray::create(Size sz, int type, int i, bool t, int f) const { create(sz.height, sz.width, type, i, t, f); }
void _OutputArray::create(int dims, const int* size, int type, int i, bool t, int f) const { CV_Assert(dims >= 1 && dims <= 2); create(size[0], dims == 2 ? size[1] : 1, type, i, t, f); }
void _OutputArray::release() const {
    if(kind_ == MAT) ((Mat*)obj)->release();
    else if(kind_ == STD_VECTOR_MAT) ((std::vector<Mat>*)obj)->clear();
}
void _OutputArray::setTo(const _InputArray& value, const _InputArray& mask) const { getMatRef().setTo(value, mask); }
void _OutputArray::assign(const Mat& m) const { getMatRef() = m; }

// ---------------------------------------------------------------------------------------------------------------------
// element-wise helpers
// ---------------------------------------------------------------------------------------------------------------------
template<typename F> static Mat cmp_scalar(const Mat& a, double s, F f) {
    CV_Assert(a.channels() == 1);
    Mat r(a.rows, a.cols, CV_8UC1);
    const size_t e = a.elemSize1(); const int dp = a.depth();
    for(int y = 0; y < a.rows; ++y) { const uchar* p = a.ptr(y); uchar* d = r.ptr(y); for(int x = 0; x < a.cols; ++x) d[x] = f(get_elem(p + e * x, dp), s) ? 255 : 0; }
    return r;
}
Mat operator==(const Mat& a, double s) { return cmp_scalar(a, s, [](double u, double v) { return u == v; }); }
Mat operator!=(const Mat& a, double s) { return cmp_scalar(a, s, [](double u, double v) { return u != v; }); }
Mat operator>(const Mat& a, double s) { return cmp_scalar(a, s, [](double u, double v) { return u > v; }); }
Mat operator<(const Mat& a, double s) { return cmp_scalar(a, s, [](double u, double v) { return u < v; }); }
Mat operator>=(const Mat& a, double s) { return cmp_scalar(a, s, [](double u, double v) { return u >= v; }); }
Mat operator<=(const Mat& a, double s) { return cmp_scalar(a, s, [](double u, double v) { return u <= v; }); }
template<typename F> static Mat cmp_mat(const Mat& a, const Mat& b, F f) {
    CV_Assert(a.rows == b.rows && a.cols == b.cols && a.type() == b.type() && a.channels() == 1);
    Mat r(a.rows, a.cols, CV_8UC1);
    const size_t e = a.elemSize1(); const int dp = a.depth();
    for(int y = 0; y < a.rows; ++y) { const uchar* p = a.ptr(y); const uchar* q = b.ptr(y); uchar* d = r.ptr(y);
        for(int x = 0; x < a.cols; ++x) d[x] = f(get_elem(p + e * x, dp), get_elem(q + e * x, dp)) ? 255 : 0; }
    return r;
}
Mat operator==(const Mat& a, const Mat& b) { return cmp_mat(a, b, [](double u, double v) { return u == v; }); }
Mat operator!=(const Mat& a, const Mat& b) { return cmp_mat(a, b, [](double u, double v) { return u != v; }); }
void compare(InputArray a_, InputArray b_, OutputArray dst, int op) {
    Mat a = a_.getMat(), b = b_.getMat(), r;
    switch(op) {
        case CMP_EQ: r = cmp_mat(a, b, [](double u, double v) { return u == v; }); break; case CMP_NE: r = cmp_mat(a, b, [](double u, double v) { return u != v; }); break;
        case CMP_GT: r = cmp_mat(a, b, [](double u, double v) { return u > v; }); break; case CMP_GE: r = cmp_mat(a, b, [](double u, double v) { return u >= v; }); break;
        case CMP_LT: r = cmp_mat(a, b, [](double u, double v) { return u < v; }); break; default: r = cmp_mat(a, b, [](double u, double v) { return u <= v; });
    }
    dst.getMatRef() = r;
}

template<typename F> static void bytewise(InputArray a_, InputArray b_, OutputArray dst, InputArray mask_, F f) {
    Mat a = a_.getMat(), b = b_.getMat(), mask = mask_.getMat();
    CV_Assert(a.rows == b.rows && a.cols == b.cols && a.type() == b.type());
    dst.create(a.rows, a.cols, a.type());
    Mat d = dst.getMat();
    const size_t esz = a.elemSize(), rowbytes = esz * a.cols;
    for(int y = 0; y < a.rows; ++y) {
        const uchar* p = a.ptr(y); const uchar* q = b.ptr(y); uchar* o = d.ptr(y);
        if(mask.empty()) for(size_t i = 0; i < rowbytes; ++i) o[i] = f(p[i], q[i]);
        else { const uchar* k = mask.ptr(y); for(size_t i = 0; i < rowbytes; ++i) if(k[i / esz]) o[i] = f(p[i], q[i]); }
    }
}
void bitwise_and(InputArray a, InputArray b, OutputArray dst, InputArray mask) { bytewise(a, b, dst, mask, [](uchar u, uchar v) { return (uchar)(u & v); }); }
void bitwise_or(InputArray a, InputArray b, OutputArray dst, InputArray mask) { bytewise(a, b, dst, mask, [](uchar u, uchar v) { return (uchar)(u | v); }); }
void bitwise_xor(InputArray a, InputArray b, OutputArray dst, InputArray mask) { bytewise(a, b, dst, mask, [](uchar u, uchar v) { return (uchar)(u ^ v); }); }
void bitwise_not(InputArray a, OutputArray dst, InputArray mask) { bytewise(a, a, dst, mask, [](uchar u, uchar) { return (uchar)~u; }); }
Mat operator&(const Mat& a, const Mat& b) { Mat r; bitwise_and(a, b, r); return r; }
Mat operator|(const Mat& a, const Mat& b) { Mat r; bitwise_or(a, b, r); return r; }
Mat operator^(const Mat& a, const Mat& b) { Mat r; bitwise_xor(a, b, r); return r; }
Mat operator~(const Mat& a) { Mat r; bitwise_not(a, r); return r; }
Mat operator/(const Mat& a, double s) { Mat r; a.convertTo(r, a.type(), 1.0 / s, 0.0); return r; }
Mat operator*(const Mat& a, double s) { Mat r; a.convertTo(r, a.type(), s, 0.0); return r; }
Mat& operator+=(Mat& a, const Mat& b) {
    CV_Assert(a.type() == b.type() && a.rows == b.rows && a.cols == b.cols);
    const int dp = a.depth(), n = a.cols * a.channels(); const size_t e = a.elemSize1();
    for(int y = 0; y < a.rows; ++y) { uchar* p = a.ptr(y); const uchar* q = b.ptr(y); for(int x = 0; x < n; ++x) set_elem(p + e * x, dp, get_elem(p + e * x, dp) + get_elem(q + e * x, dp)); }
    return a;
}
Mat Mat::mul(const Mat& m, double scale) const {
    CV_Assert(type() == m.type() && rows == m.rows && cols == m.cols);
    Mat r(rows, cols, type());
    const int dp = depth(), n = cols * channels(); const size_t e = elemSize1();
    for(int y = 0; y < rows; ++y) { const uchar* p = ptr(y); const uchar* q = m.ptr(y); uchar* o = r.ptr(y); for(int x = 0; x < n; ++x) set_elem(o + e * x, dp, get_elem(p + e * x, dp) * get_elem(q + e * x, dp) * scale); }
    return r;
}
double kmeans(InputArray, int, InputOutputArray, TermCriteria, int, int, OutputArray) { unsupported("kmeans"); return 0; }
template<typename F> static Mat& bits_scalar(Mat& a, const Scalar& s, F f) {
    CV_Assert(a.depth() == CV_8U);
    const int cn = a.channels();
    uchar v[4]; for(int c = 0; c < 4; ++c) v[c] = saturate_cast<uchar>(s.val[c]);
    for(int y = 0; y < a.rows; ++y) { uchar* p = a.ptr(y); for(int x = 0; x < a.cols; ++x) for(int c = 0; c < cn; ++c) p[x * cn + c] = f(p[x * cn + c], v[c & 3]); }
    return a;
}
Mat& operator|=(Mat& a, const Scalar& s) { return bits_scalar(a, s, [](uchar u, uchar v) { return (uchar)(u | v); }); }
Mat& operator&=(Mat& a, const Scalar& s) { return bits_scalar(a, s, [](uchar u, uchar v) { return (uchar)(u & v); }); }

int countNonZero(InputArray a_) {
    Mat a = a_.getMat();
    CV_Assert(a.channels() == 1);
    int n = 0; const size_t e = a.elemSize1(); const int dp = a.depth();
    for(int y = 0; y < a.rows; ++y) { const uchar* p = a.ptr(y); for(int x = 0; x < a.cols; ++x) n += get_elem(p + e * x, dp) != 0; }
    return n;
}
/// cv::sum accumulates every channel in double, in raster order
Scalar sum(InputArray a_) {
    Mat a = a_.getMat();
    Scalar s; const int cn = a.channels(), dp = a.depth(); const size_t e = a.elemSize1();
    CV_Assert(cn <= 4);
    for(int y = 0; y < a.rows; ++y) { const uchar* p = a.ptr(y); for(int x = 0; x < a.cols; ++x) for(int c = 0; c < cn; ++c) s.val[c] += get_elem(p + e * ((size_t)x * cn + c), dp); }
    return s;
}
Scalar mean(InputArray a_, InputArray mask) {
    CV_Assert(mask.empty());
    Mat a = a_.getMat(); Scalar s = sum(a); const double n = (double)a.total();
    for(int c = 0; c < 4; ++c) s.val[c] = n ? s.val[c] / n : 0;
    return s;
}
void split(const Mat& src, Mat* mv) {
    const int cn = src.channels(); const size_t e = src.elemSize1();
    for(int c = 0; c < cn; ++c) mv[c].create(src.rows, src.cols, src.depth());
    for(int y = 0; y < src.rows; ++y) { const uchar* p = src.ptr(y);
        for(int c = 0; c < cn; ++c) { uchar* d = mv[c].ptr(y); for(int x = 0; x < src.cols; ++x) std::memcpy(d + e * x, p + e * ((size_t)x * cn + c), e); } }
}
void split(InputArray src_, OutputArrayOfArrays mv) {
    Mat src = src_.getMat();
    CV_Assert(mv.kind() == _InputArray::STD_VECTOR_MAT);
    std::vector<Mat>& v = *(std::vector<Mat>*)mv.getObj();
    v.resize((size_t)src.channels());
    split(src, v.data());
}
void merge(const Mat* mv, size_t count, OutputArray dst) {
    CV_Assert(count >= 1);
    const int cn = (int)count; const size_t e = mv[0].elemSize1();
    Mat d(mv[0].rows, mv[0].cols, CV_MAKETYPE(mv[0].depth(), cn));
    for(int y = 0; y < d.rows; ++y) { uchar* o = d.ptr(y);
        for(int c = 0; c < cn; ++c) { CV_Assert(mv[c].channels() == 1); const uchar* p = mv[c].ptr(y); for(int x = 0; x < d.cols; ++x) std::memcpy(o + e * ((size_t)x * cn + c), p + e * x, e); } }
    dst.getMatRef() = d;
}
void merge(InputArrayOfArrays mv, OutputArray dst) { std::vector<Mat> v; mv.getMatVector(v); merge(v.data(), v.size(), dst); }
template<typename F> static void minmax_impl(const Mat& a, const Mat& b, Mat& dst, F f) {
    CV_Assert(a.rows == b.rows && a.cols == b.cols && a.type() == b.type());
    Mat d; if(dst.data == a.data || dst.data == b.data) d = dst; else { dst.create(a.rows, a.cols, a.type()); d = dst; }
    const int n = a.cols * a.channels(), dp = a.depth(); const size_t e = a.elemSize1();
    for(int y = 0; y < a.rows; ++y) { const uchar* p = a.ptr(y); const uchar* q = b.ptr(y); uchar* o = d.ptr(y);
        for(int x = 0; x < n; ++x) { const double u = get_elem(p + e * x, dp), v = get_elem(q + e * x, dp); set_elem(o + e * x, dp, f(u, v)); } }
}
void max(const Mat& a, const Mat& b, Mat& dst) { minmax_impl(a, b, dst, [](double u, double v) { return u > v ? u : v; }); }
void min(const Mat& a, const Mat& b, Mat& dst) { minmax_impl(a, b, dst, [](double u, double v) { return u < v ? u : v; }); }
void max(InputArray a, InputArray b, OutputArray dst) { Mat A = a.getMat(), B = b.getMat(); max(A, B, dst.getMatRef()); }
void min(InputArray a, InputArray b, OutputArray dst) { Mat A = a.getMat(), B = b.getMat(); min(A, B, dst.getMatRef()); }
void absdiff(InputArray a_, InputArray b_, OutputArray dst) {
    Mat a = a_.getMat(), b = b_.getMat();
    minmax_impl(a, b, dst.getMatRef(), [](double u, double v) { return std::fabs(u - v); });
}
void minMaxIdx(InputArray src_, double* minVal, double* maxVal, int* minIdx, int* maxIdx, InputArray mask) {
    CV_Assert(mask.empty());
    Mat a = src_.getMat();
    CV_Assert(a.channels() == 1 && !a.empty());
    double mn = DBL_MAX, mx = -DBL_MAX; int mny = 0, mnx = 0, mxy = 0, mxx = 0;
    const size_t e = a.elemSize1(); const int dp = a.depth();
    for(int y = 0; y < a.rows; ++y) { const uchar* p = a.ptr(y); for(int x = 0; x < a.cols; ++x) { const double v = get_elem(p + e * x, dp);
        if(v < mn) { mn = v; mny = y; mnx = x; } if(v > mx) { mx = v; mxy = y; mxx = x; } } }
    if(minVal) *minVal = mn; if(maxVal) *maxVal = mx;
    if(minIdx) { minIdx[0] = mny; minIdx[1] = mnx; } if(maxIdx) { maxIdx[0] = mxy; maxIdx[1] = mxx; }
}
void minMaxLoc(InputArray src, double* minVal, double* maxVal, Point* minLoc, Point* maxLoc, InputArray mask) {
    int a[2], b[2]; minMaxIdx(src, minVal, maxVal, a, b, mask);
    if(minLoc) *minLoc = Point(a[1], a[0]); if(maxLoc) *maxLoc = Point(b[1], b[0]);
}
double norm(InputArray a_, int normType, InputArray mask) {
    CV_Assert(mask.empty());
    Mat a = a_.getMat(); double s = 0; const int n = a.cols * a.channels(), dp = a.depth(); const size_t e = a.elemSize1();
    for(int y = 0; y < a.rows; ++y) { const uchar* p = a.ptr(y); for(int x = 0; x < n; ++x) { const double v = get_elem(p + e * x, dp);
        if(normType == NORM_L1) s += std::fabs(v); else if(normType == NORM_INF) s = std::max(s, std::fabs(v)); else s += v * v; } }
    return (normType == NORM_L2) ? std::sqrt(s) : s;
}
double norm(InputArray, InputArray, int, InputArray) { unsupported("norm(a,b)"); return 0; }
double determinant(InputArray) { unsupported("determinant"); return 0; }
double invert(InputArray, OutputArray, int) { unsupported("invert"); return 0; }
void copyMakeBorder(InputArray, OutputArray, int, int, int, int, int, const Scalar&) { unsupported("copyMakeBorder"); }
void normalize(InputArray src_, InputOutputArray dst, double alpha, double beta, int norm_type, int dtype, InputArray mask) {
    // reached from the reference's debug displays and from EdgeDetectorLBSP::apply (bNormalizeOutput); NORM_MINMAX to [alpha,beta]
    CV_Assert(mask.empty() && norm_type == NORM_MINMAX);
    Mat src = src_.getMat(); double mn, mx; minMaxIdx(src.reshape(1), &mn, &mx);
    if(src.type() == CV_8UC1 && dtype < 0 && std::min(alpha, beta) == 0 && std::max(alpha, beta) == 255) {
        // OpenCV converts in FLOAT here (multiply, add, round half to even): the oracle's restatement, pinned against cv2 4.13
        Mat out = src.clone(); CV_Assert(out.isContinuous());
        lvo::EdgeDetectorLBSP::normalize_minmax_u8(out.data, (size_t)out.rows * out.cols);
        Mat& d = dst.getMatRef();   // like convertTo: an output of the right size and type is written in place (the caller may hold another header on it)
        if(d.data && d.rows == out.rows && d.cols == out.cols && d.type() == out.type()) { for(int y = 0; y < out.rows; ++y) std::memcpy(d.ptr(y), out.ptr(y), (size_t)out.cols); }
        else d = out;
        return;
    }
    const double lo = std::min(alpha, beta), hi = std::max(alpha, beta), sc = mx > mn ? (hi - lo) / (mx - mn) : 0.0;
    Mat out; src.convertTo(out, dtype < 0 ? src.type() : dtype, sc, lo - mn * sc); dst.getMatRef() = out;
}

/// cv::addWeighted(src1, alpha, src2, beta, gamma, dst, dtype). SURVEY Appendix E (probed with cv2 4.13): for f32 + u8 -> f32 the
/// result is f32( f64(a)*alpha + f64(b)*beta + gamma ): accumulated in double, rounded once.
void addWeighted(InputArray a_, double alpha, InputArray b_, double beta, double gamma, OutputArray dst, int dtype) {
    Mat a = a_.getMat(), b = b_.getMat();
    CV_Assert(a.rows == b.rows && a.cols == b.cols && a.channels() == b.channels());
    const int ddepth = dtype < 0 ? a.depth() : CV_MAT_DEPTH(dtype), cn = a.channels();
    Mat d;
    if(dst.getObj() == a_.getObj() && ddepth == a.depth()) d = a; else d.create(a.rows, a.cols, CV_MAKETYPE(ddepth, cn));
    const size_t ea = a.elemSize1(), eb = b.elemSize1(), ed = d.elemSize1();
    for(int y = 0; y < a.rows; ++y) { const uchar* p = a.ptr(y); const uchar* q = b.ptr(y); uchar* o = d.ptr(y);
        for(int x = 0; x < a.cols * cn; ++x) set_elem(o + ed * x, ddepth, get_elem(p + ea * x, a.depth()) * alpha + get_elem(q + eb * x, b.depth()) * beta + gamma); }
    dst.getMatRef() = d;
}

// ---------------------------------------------------------------------------------------------------------------------
// imgproc
// ---------------------------------------------------------------------------------------------------------------------
Mat getStructuringElement(int shape, Size ksize, Point) { CV_Assert(shape == MORPH_RECT); return Mat(ksize, CV_8UC1, Scalar(1)); }

static Mat continuous_u8(const Mat& m) { if(m.isContinuous()) return m; return m.clone(); }
static int rect_radius(InputArray kernel, int iterations) {
    // default (empty) kernel = 3x3 rect; `iterations` applications of a (2r+1)^2 rect equal one of a (2*r*iterations+1)^2 rect when
    // pixels outside the image are ignored (the default morphology border)
    Mat k = kernel.getMat(); int r = 1;
    if(!k.empty()) { CV_Assert(k.rows == k.cols && (k.rows & 1) && countNonZero(k) == k.rows * k.cols); r = k.rows / 2; }
    return r * iterations;
}
static void morph(InputArray src_, OutputArray dst, InputArray kernel, int iterations, bool dil, int borderType, const Scalar& bv) {
    Mat src = continuous_u8(src_.getMat());
    CV_Assert(src.type() == CV_8UC1 && borderType == BORDER_CONSTANT && bv == morphologyDefaultBorderValue());
    Mat out(src.rows, src.cols, CV_8UC1);
    lvo::morph_rect(src.data, out.data, src.cols, src.rows, rect_radius(kernel, iterations), dil);
    dst.create(src.rows, src.cols, CV_8UC1);
    out.copyTo(dst.getMatRef());
}
void erode(InputArray s, OutputArray d, InputArray k, Point, int it, int bt, const Scalar& bv) { morph(s, d, k, it, false, bt, bv); }
void dilate(InputArray s, OutputArray d, InputArray k, Point, int it, int bt, const Scalar& bv) { morph(s, d, k, it, true, bt, bv); }
void morphologyEx(InputArray s, OutputArray d, int op, InputArray k, Point a, int it, int bt, const Scalar& bv) {
    Mat tmp;
    if(op == MORPH_CLOSE) { dilate(s, tmp, k, a, it, bt, bv); erode(tmp, d, k, a, it, bt, bv); }
    else if(op == MORPH_OPEN) { erode(s, tmp, k, a, it, bt, bv); dilate(tmp, d, k, a, it, bt, bv); }
    else unsupported("morphologyEx op");
}
static bool is_binary(const Mat& m) { for(int y = 0; y < m.rows; ++y) { const uchar* p = m.ptr(y); for(int x = 0; x < m.cols; ++x) if(p[x] != 0 && p[x] != 255) return false; } return true; }
void medianBlur(InputArray src_, OutputArray dst, int ksize) {
    Mat src = continuous_u8(src_.getMat());
    CV_Assert(src.type() == CV_8UC1 && (ksize & 1) && ksize >= 3);
    if(!is_binary(src)) unsupported("medianBlur on a non-binary image (the hot path only filters {0,255} masks)");
    Mat out(src.rows, src.cols, CV_8UC1);
    lvo::median_binary(src.data, out.data, src.cols, src.rows, ksize);
    dst.create(src.rows, src.cols, CV_8UC1);
    out.copyTo(dst.getMatRef());
}
int floodFill(InputOutputArray image, Point seed, Scalar newVal, Rect*, Scalar lo, Scalar up, int flags) {
    Mat& m = image.getMatRef();
    CV_Assert(m.type() == CV_8UC1 && m.isContinuous() && seed == Point(0, 0) && newVal.val[0] == 255 && lo == Scalar() && up == Scalar() && (flags & 0xFF) == 4);
    if(!is_binary(m)) unsupported("floodFill on a non-binary image");
    lvo::floodfill_from_origin(m.data, m.cols, m.rows);
    return 0;
}
void resize(InputArray src_, OutputArray dst, Size dsize, double fx, double fy, int interpolation) {
    Mat src = src_.getMat();
    if(dsize.width <= 0 || dsize.height <= 0) dsize = Size(saturate_cast<int>(src.cols * fx), saturate_cast<int>(src.rows * fy));
    CV_Assert(!src.empty() && dsize.width > 0 && dsize.height > 0);
    Mat out(dsize, src.type());
    if(interpolation == INTER_NEAREST) {
        // OpenCV: sx = min(floor(x * (1/fx)), cols-1) with the inverse scale in double
        const double ifx = (double)src.cols / dsize.width, ify = (double)src.rows / dsize.height; const size_t esz = src.elemSize();
        for(int y = 0; y < dsize.height; ++y) { const int sy = std::min((int)std::floor(y * ify), src.rows - 1);
            for(int x = 0; x < dsize.width; ++x) { const int sx = std::min((int)std::floor(x * ifx), src.cols - 1); std::memcpy(out.ptr(y) + esz * x, src.ptr(sy) + esz * sx, esz); } }
    } else if(interpolation == INTER_AREA && src.depth() == CV_8U && dsize.width <= src.cols && dsize.height <= src.rows) {
        Mat s = continuous_u8(src);
        const int cn = src.channels();
        if(src.cols % dsize.width == 0 && src.rows % dsize.height == 0 && src.cols / dsize.width == src.rows / dsize.height)
            lvo::resize_area_exact(s.data, s.cols, s.rows, cn, s.cols / dsize.width, out.data);
        else lvo::resize_area_general(s.data, s.cols, s.rows, cn, dsize.width, dsize.height, out.data);
    } else unsupported("resize mode (only INTER_NEAREST and shrinking 8-bit INTER_AREA are on the hot path)");
    dst.getMatRef() = out;
}
/// cv::blur(32F, 3x3): SURVEY Appendix E / oracle lvo_pawcs.hpp blur3: row sums then column sums in double, x 1/9, one rounding
void blur(InputArray src_, OutputArray dst, Size ksize, Point, int borderType) {
    Mat src = src_.getMat();
    CV_Assert(src.type() == CV_32FC1 && ksize == Size(3, 3) && borderType == BORDER_REPLICATE);
    const int W = src.cols, H = src.rows;
    std::vector<double> rs((size_t)W * H);
    for(int y = 0; y < H; ++y) { const float* r = src.ptr<float>(y);
        for(int x = 0; x < W; ++x) rs[(size_t)y * W + x] = (double)r[std::max(x - 1, 0)] + (double)r[x] + (double)r[std::min(x + 1, W - 1)]; }
    Mat out(H, W, CV_32FC1);
    for(int y = 0; y < H; ++y) for(int x = 0; x < W; ++x) {
        const double s = rs[(size_t)std::max(y - 1, 0) * W + x] + rs[(size_t)y * W + x] + rs[(size_t)std::min(y + 1, H - 1) * W + x];
        out.at<float>(y, x) = (float)(s * (1.0 / 9.0));
    }
    dst.getMatRef() = out;
}
static inline int reflect101(int p, int len) { if(len == 1) return 0; while(p < 0 || p >= len) { if(p < 0) p = -p; if(p >= len) p = 2 * len - 2 - p; } return p; }
/// cv::GaussianBlur(8-bit, 3x3, sigma 0, BORDER_DEFAULT) as PBAS calls it (PBAS.cpp:83, :117): kernel [1 2 1]/4 per axis, OpenCV's fixed-point
/// path for 8-bit images (exact weighted sum, + 8, >> 4). The chain blur -> Scharr -> convertScaleAbs -> addWeighted(0.5, 0.5) is pinned
/// bit-exactly against cv2 4.13 as the oracle's pbas_gradient_image (tests/test_pbas_oracle_cpu.py); the stages here use the same arithmetic.
void GaussianBlur(InputArray src_, OutputArray dst, Size ksize, double sx, double sy, int borderType) {
    Mat src = continuous_u8(src_.getMat());
    CV_Assert(src.depth() == CV_8U && ksize == Size(3, 3) && sx == 0 && sy == 0 && borderType == BORDER_DEFAULT);
    const int W = src.cols, H = src.rows, C = src.channels();
    Mat out(H, W, src.type());
    static const int w3[3] = {1, 2, 1};
    for(int y = 0; y < H; ++y) for(int x = 0; x < W; ++x) for(int c = 0; c < C; ++c) {
        int s = 0;
        for(int dy = -1; dy <= 1; ++dy) for(int dx = -1; dx <= 1; ++dx) s += w3[dy + 1] * w3[dx + 1] * src.ptr(reflect101(y + dy, H))[reflect101(x + dx, W) * C + c];
        out.ptr(y)[x * C + c] = (uchar)((s + 8) >> 4);
    }
    dst.getMatRef() = out;
}
/// cv::Scharr(8-bit -> 16S, (dx,dy) = (1,0) or (0,1), scale 1, delta 0, BORDER_DEFAULT): [-3 0 3; -10 0 10; -3 0 3] and its transpose
void Scharr(InputArray src_, OutputArray dst, int ddepth, int dx, int dy, double scale, double delta, int borderType) {
    Mat src = continuous_u8(src_.getMat());
    CV_Assert(src.depth() == CV_8U && ddepth == CV_16S && scale == 1 && delta == 0 && borderType == BORDER_DEFAULT && dx + dy == 1 && dx >= 0 && dy >= 0);
    const int W = src.cols, H = src.rows, C = src.channels();
    Mat out(H, W, CV_MAKETYPE(CV_16S, C));
    auto B = [&](int y, int x, int c) { return (int)src.ptr(reflect101(y, H))[reflect101(x, W) * C + c]; };
    for(int y = 0; y < H; ++y) for(int x = 0; x < W; ++x) for(int c = 0; c < C; ++c) {
        const int g = dx ? 3 * (B(y - 1, x + 1, c) - B(y - 1, x - 1, c)) + 10 * (B(y, x + 1, c) - B(y, x - 1, c)) + 3 * (B(y + 1, x + 1, c) - B(y + 1, x - 1, c))
                         : 3 * (B(y + 1, x - 1, c) - B(y - 1, x - 1, c)) + 10 * (B(y + 1, x, c) - B(y - 1, x, c)) + 3 * (B(y + 1, x + 1, c) - B(y - 1, x + 1, c));
        out.ptr<short>(y)[x * C + c] = (short)g;
    }
    dst.getMatRef() = out;
}
/// cv::convertScaleAbs(16S -> 8U, alpha 1, beta 0): saturate(|v|)
void convertScaleAbs(InputArray src_, OutputArray dst, double alpha, double beta) {
    Mat src = src_.getMat();
    CV_Assert(src.depth() == CV_16S && alpha == 1 && beta == 0);
    Mat out(src.rows, src.cols, CV_MAKETYPE(CV_8U, src.channels()));
    const int n = src.cols * src.channels();
    for(int y = 0; y < src.rows; ++y) { const short* p = src.ptr<short>(y); uchar* o = out.ptr(y); for(int x = 0; x < n; ++x) o[x] = (uchar)std::min(std::abs((int)p[x]), 255); }
    dst.getMatRef() = out;
}
/// cv::accumulateWeighted(u8 -> f32): dst = src*alpha + dst*(1-alpha), in float with separate multiplies and add (Appendix E)
void accumulateWeighted(InputArray src_, InputOutputArray dst_, double alpha, InputArray mask) {
    CV_Assert(mask.empty());
    Mat src = src_.getMat(); Mat& dst = dst_.getMatRef();
    CV_Assert(src.rows == dst.rows && src.cols == dst.cols && src.channels() == dst.channels() && dst.depth() == CV_32F && (src.depth() == CV_8U || src.depth() == CV_32F));
    const float a = (float)alpha, b = 1.0f - a; const int n = src.cols * src.channels();
    for(int y = 0; y < src.rows; ++y) { float* d = dst.ptr<float>(y);
        for(int x = 0; x < n; ++x) { const float sv = src.depth() == CV_8U ? (float)src.ptr(y)[x] : src.ptr<float>(y)[x]; const float s = sv * a, t = d[x] * b; d[x] = s + t; } }
}
/// cv::accumulateProduct(f32, f32, f32, mask): dst += src1*src2 where mask != 0, float multiply then float add
void accumulateProduct(InputArray s1_, InputArray s2_, InputOutputArray dst_, InputArray mask_) {
    Mat s1 = s1_.getMat(), s2 = s2_.getMat(), mask = mask_.getMat(); Mat& dst = dst_.getMatRef();
    CV_Assert(s1.type() == CV_32FC1 && s2.type() == CV_32FC1 && dst.type() == CV_32FC1 && s1.size() == s2.size() && s1.size() == dst.size());
    for(int y = 0; y < s1.rows; ++y) { const float* p = s1.ptr<float>(y); const float* q = s2.ptr<float>(y); float* d = dst.ptr<float>(y); const uchar* k = mask.empty() ? nullptr : mask.ptr(y);
        for(int x = 0; x < s1.cols; ++x) if(!k || k[x]) { const float t = p[x] * q[x]; d[x] += t; } }
}
/// cv::cvtColor(COLOR_GRAY2BGR) (ViBe.cpp / PBAS.cpp: gray frames into the 3-channel model): the value replicated
void cvtColor(InputArray src_, OutputArray dst, int code, int) {
    Mat src = src_.getMat();
    if(code != COLOR_GRAY2BGR || src.type() != CV_8UC1) unsupported("cvtColor");
    Mat out(src.rows, src.cols, CV_8UC3);
    for(int y = 0; y < src.rows; ++y) { const uchar* p = src.ptr(y); uchar* o = out.ptr(y); for(int x = 0; x < src.cols; ++x) o[3 * x] = o[3 * x + 1] = o[3 * x + 2] = p[x]; }
    dst.getMatRef() = out;
}
void circle(InputOutputArray, Point, int, const Scalar&, int, int, int) {}
void putText(InputOutputArray, const String&, Point, int, double, Scalar, int, int, bool) {}
void rectangle(InputOutputArray, Rect, const Scalar&, int, int, int) {}
void line(InputOutputArray, Point, Point, const Scalar&, int, int, int) {}

// ---- highgui (debug displays): no-ops
void imshow(const String&, InputArray) {}
int waitKey(int) { return -1; }
void namedWindow(const String&, int) {}
void destroyWindow(const String&) {}
void destroyAllWindows() {}
void moveWindow(const String&, int, int) {}
void resizeWindow(const String&, int, int) {}
void setMouseCallback(const String&, MouseCallback, void*) {}
bool imwrite(const String&, InputArray, const std::vector<int>&) { return false; }
Mat imread(const String&, int) { return Mat(); }

// ---- features2d
void KeyPointsFilter::runByImageBorder(std::vector<KeyPoint>& kps, Size sz, int border) {
    if(border <= 0) return;
    if(sz.height <= border * 2 || sz.width <= border * 2) { kps.clear(); return; }
    const Rect2f r((float)border, (float)border, (float)(sz.width - 2 * border), (float)(sz.height - 2 * border));
    kps.erase(std::remove_if(kps.begin(), kps.end(), [&](const KeyPoint& k) { return !r.contains(k.pt); }), kps.end());
}
void KeyPointsFilter::runByPixelsMask(std::vector<KeyPoint>& kps, const Mat& mask) {
    if(mask.empty()) return;
    kps.erase(std::remove_if(kps.begin(), kps.end(), [&](const KeyPoint& k) { return mask.at<uchar>((int)(k.pt.y + 0.5f), (int)(k.pt.x + 0.5f)) == 0; }), kps.end());
}
void Feature2D::detect(InputArray image, std::vector<KeyPoint>& keypoints, InputArray mask) { Mat none; detectAndCompute(image, mask, keypoints, noArray(), false); }
void Feature2D::detect(InputArrayOfArrays, std::vector<std::vector<KeyPoint>>&, InputArrayOfArrays) { unsupported("Feature2D::detect(collection)"); }
void Feature2D::compute(InputArray image, std::vector<KeyPoint>& keypoints, OutputArray descriptors) { detectAndCompute(image, noArray(), keypoints, descriptors, true); }
void Feature2D::compute(InputArrayOfArrays, std::vector<std::vector<KeyPoint>>&, OutputArrayOfArrays) { unsupported("Feature2D::compute(collection)"); }
void Feature2D::detectAndCompute(InputArray, InputArray, std::vector<KeyPoint>&, OutputArray, bool) { unsupported("Feature2D::detectAndCompute"); }

} // namespace cv
