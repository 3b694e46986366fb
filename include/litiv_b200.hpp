// litiv_b200 — header-only C++ drop-in classes over the C ABI (include/litiv_b200.h).
//
// Same class and method names as the reference (modules/video/include/litiv/video/BackgroundSubtractionUtils.hpp:24-47,
// BackgroundSubtractorSuBSENSE.hpp:49-133, BackgroundSubtractorLOBSTER.hpp:133-153): initialize(img, ROI),
// apply(img, fgmask, learningRate), getBackgroundImage, getBackgroundDescriptorsImage, refreshModel, setROI, getROICopy,
// setAutomaticModelReset, getDefaultLearningRate. Failures throw lv::Exception-like std::runtime_error carrying the
// reference's assertion text (the reference throws lv::Exception : std::runtime_error, utils/cxx.hpp:189-204).
//
// With OpenCV available (define LITIV_B200_WITH_OPENCV before including) the classes derive from cv::BackgroundSubtractor
// and take cv::Mat / cv::InputArray exactly like the reference; without it they take the plain lvb::ImageView below.
#pragma once
#include "litiv_b200.h"
#include <cmath>
#include <cstdint>
#include <limits>
#include <stdexcept>
#include <string>
#include <vector>
#ifdef LITIV_B200_WITH_OPENCV
#include <opencv2/core.hpp>
#include <opencv2/video/background_segm.hpp>
#endif

namespace lvb {

struct Exception : std::runtime_error { using std::runtime_error::runtime_error; };
inline void check(int rc) { if(rc != 0) throw Exception(lvb_last_error()); }

/// minimal continuous 8-bit image view {data, rows, cols, channels, step}
struct ImageView {
    const uint8_t* data = nullptr; int rows = 0, cols = 0, channels = 0; size_t step = 0;
    ImageView() {}
    ImageView(const uint8_t* d, int r, int c, int ch, size_t s = 0) : data(d), rows(r), cols(c), channels(ch), step(s ? s : (size_t)c * ch) {}
    bool empty() const { return !data || rows <= 0 || cols <= 0; }
    bool isContinuous() const { return step == (size_t)cols * channels; }
#ifdef LITIV_B200_WITH_OPENCV
    ImageView(const cv::Mat& m) : data(m.data), rows(m.rows), cols(m.cols), channels(m.channels()), step(m.step.p[0]) {
        if(!m.empty() && m.depth() != CV_8U) throw Exception("input image type/size mismatch with initialization type/size");
    }
#endif
};

#ifdef LITIV_B200_WITH_OPENCV
struct SubtractorBase : public cv::BackgroundSubtractor {
#else
struct SubtractorBase {
#endif
    virtual ~SubtractorBase() { if(m_h) lvb_destroy(m_h); }
    SubtractorBase(const SubtractorBase&) = delete;
    SubtractorBase& operator=(const SubtractorBase&) = delete;

    /// IIBackgroundSubtractor::initialize(img) / initialize(img, ROI)
    void initialize(const ImageView& img) { initialize(img, ImageView()); }
    virtual void initialize(const ImageView& img, const ImageView& roi) {
        if(img.empty()) throw Exception("provided image for initialization must be non-empty, continuous, and of type 8UC1/3/4");
        if(!roi.empty() && (roi.rows != img.rows || roi.cols != img.cols || roi.channels != 1 || !roi.isContinuous()))
            throw Exception("provided ROI mat size must be equal to the init frame size, and its type must be 8UC1");
        check(lvb_initialize(m_h, img.data, img.cols, img.rows, img.channels, img.step, roi.empty() ? nullptr : roi.data));
        m_rows = img.rows; m_cols = img.cols; m_channels = img.channels;
    }
    /// IBackgroundSubtractor::apply(img, fgmask, learningRate); fgmask must hold rows*cols bytes
    virtual void apply(const ImageView& img, uint8_t* fgmask, double learningRate) {
        if(img.rows != m_rows || img.cols != m_cols || img.channels != m_channels) throw Exception(m_rows ? "input image type/size mismatch with initialization type/size" : "algo & model must be initialized first");
        if(!img.isContinuous()) throw Exception("input image data must be continuous");
        check(lvb_apply(m_h, img.data, fgmask, learningRate));
    }
    void apply(const ImageView& img, uint8_t* fgmask) { apply(img, fgmask, getDefaultLearningRate()); }
    void apply(const ImageView& img, std::vector<uint8_t>& fgmask, double learningRate) { fgmask.resize((size_t)m_rows * m_cols); apply(img, fgmask.data(), learningRate); }
    /// asynchronous pair (the `apply_cuda` async mode sketched in apps/changedet/src/main.cpp:274-282)
    void apply_async(const ImageView& img, uint8_t* fgmask, double learningRate) {
        if(img.rows != m_rows || img.cols != m_cols || img.channels != m_channels) throw Exception(m_rows ? "input image type/size mismatch with initialization type/size" : "algo & model must be initialized first");
        if(!img.isContinuous()) throw Exception("input image data must be continuous");
        if(!fgmask) throw Exception("output mask must be provided");
        check(lvb_apply_async(m_h, img.data, fgmask, learningRate));
    }
    void sync_next() { check(lvb_sync_next(m_h)); }
    void sync() { check(lvb_sync(m_h)); }
    /// order later work on the instance's CUDA stream behind the side-stream work of the frames enqueued so far
    void flush() { check(lvb_flush(m_h)); }
    /// device-resident frame (what a cv::cuda::GpuMat overload binds to)
    void apply_device(const uint8_t* d_img, size_t d_step, uint8_t* d_fgmask, double learningRate) { check(lvb_apply_device(m_h, d_img, d_step, d_fgmask, learningRate)); }

    virtual void getBackgroundImage(uint8_t* out) const { check(lvb_get_background_image(m_h, out)); }
    /// getBackgroundImage into device memory (rows*cols*channels bytes): what the cv::cuda::GpuMat overload of apps/changedet/src/main.cpp:304-307 binds to
    void getBackgroundImageDevice(uint8_t* d_out) const { check(lvb_get_background_image_device(m_h, d_out)); }
    virtual void getBackgroundDescriptorsImage(uint16_t* out) const { check(lvb_get_background_descriptors_image(m_h, out)); }
    virtual double getDefaultLearningRate() const { return lvb_default_learning_rate(m_algo); }
    virtual void setAutomaticModelReset(bool b) { check(lvb_set_auto_model_reset(m_h, b ? 1 : 0)); }
    /// IIBackgroundSubtractor::validateROI (BackgroundSubtractionUtils.cpp:28-36): clears the LBSP border (2 px) of a writable ROI in place
    virtual void validateROI(uint8_t* roi, int rows, int cols) const {
        if(!roi || rows <= 0 || cols <= 0) throw Exception("provided ROI must be non-empty and of type 8UC1");
        check(lvb_validate_roi(roi, cols, rows, 2));
    }
    virtual void setROI(const ImageView& roi) {
        if(roi.empty() || roi.channels != 1 || !roi.isContinuous()) throw Exception("provided ROI must be non-empty and of type 8UC1");
        if(m_rows && (roi.rows != m_rows || roi.cols != m_cols)) throw Exception("provided ROI mat size must be equal to the init frame size, and its type must be 8UC1");
        check(lvb_set_roi(m_h, roi.data));
    }
    virtual std::vector<uint8_t> getROICopy() const { std::vector<uint8_t> r((size_t)m_rows * m_cols); check(lvb_get_roi(m_h, r.data())); return r; }
    void refreshModel(float fSamplesRefreshFrac, bool bForceFGUpdate = false) { check(lvb_refresh_model(m_h, fSamplesRefreshFrac, bForceFGUpdate ? 1 : 0)); }
    lvb_handle handle() const { return m_h; }

#ifdef LITIV_B200_WITH_OPENCV
    // cv::BackgroundSubtractor interface, as in the reference
    void initialize(const cv::Mat& img, const cv::Mat& roi) { initialize(ImageView(img), roi.empty() ? ImageView() : ImageView(roi)); }
    void apply(cv::InputArray image, cv::OutputArray fgmask, double learningRate) override {
        cv::Mat img = image.getMat();
        fgmask.create(img.size(), CV_8UC1);
        cv::Mat m = fgmask.getMat();
        apply(ImageView(img), m.data, learningRate);
    }
    void getBackgroundImage(cv::OutputArray out) const override {
        out.create(m_rows, m_cols, CV_8UC(m_channels));
        cv::Mat m = out.getMat();
        getBackgroundImage(m.data);
    }
    void validateROI(cv::Mat& roi) const {
        if(roi.empty() || roi.type() != CV_8UC1 || !roi.isContinuous()) throw Exception("provided ROI must be non-empty and of type 8UC1");
        validateROI(roi.data, roi.rows, roi.cols);
    }
    void setROI(cv::Mat& roi) { validateROI(roi); setROI(ImageView(roi)); }
    cv::Mat getROICopyMat() const { std::vector<uint8_t> r = getROICopy(); cv::Mat m(m_rows, m_cols, CV_8UC1); std::memcpy(m.data, r.data(), r.size()); return m; }
#endif
protected:
    SubtractorBase(int algo, const lvb_params& p, int device, uint64_t seed) : m_algo(algo) { check(lvb_create(algo, &p, device, seed, &m_h)); }
    static lvb_params defaults(int algo) { lvb_params p; check(lvb_default_params(algo, &p)); return p; }
    lvb_handle m_h = nullptr;
    int m_algo, m_rows = 0, m_cols = 0, m_channels = 0;
};

/// BackgroundSubtractorSuBSENSE_<lv::CUDA> (ctor arguments and defaults: BackgroundSubtractorSuBSENSE.hpp:52-57)
struct BackgroundSubtractorSuBSENSE : SubtractorBase {
    explicit BackgroundSubtractorSuBSENSE(size_t nDescDistThresholdOffset = 3, size_t nMinColorDistThreshold = 30, size_t nBGSamples = 50,
                                          size_t nRequiredBGSamples = 2, size_t nSamplesForMovingAvgs = 100, float fRelLBSPThreshold = 0.333f,
                                          int device = 0, uint64_t seed = 0)
        : SubtractorBase(LVB_ALGO_SUBSENSE, make(nDescDistThresholdOffset, nMinColorDistThreshold, nBGSamples, nRequiredBGSamples, nSamplesForMovingAvgs, fRelLBSPThreshold), device, seed) {}
private:
    static lvb_params make(size_t d, size_t c, size_t n, size_t r, size_t a, float rel) {
        lvb_params p = defaults(LVB_ALGO_SUBSENSE);
        p.desc_dist_threshold = (int32_t)d; p.color_dist_threshold = (int32_t)c; p.n_samples = (int32_t)n; p.n_required = (int32_t)r;
        p.n_samples_for_moving_avgs = (int32_t)a; p.rel_lbsp_threshold = rel;
        return p;
    }
};

/// BackgroundSubtractorLOBSTER_<lv::CUDA> (ctor arguments and defaults: BackgroundSubtractorLOBSTER.hpp:50-55)
struct BackgroundSubtractorLOBSTER : SubtractorBase {
    explicit BackgroundSubtractorLOBSTER(size_t nDescDistThreshold = 4, size_t nColorDistThreshold = 30, size_t nBGSamples = 35,
                                         size_t nRequiredBGSamples = 2, size_t nLBSPThresholdOffset = 0, float fRelLBSPThreshold = 0.333f,
                                         int device = 0, uint64_t seed = 0)
        : SubtractorBase(LVB_ALGO_LOBSTER, make(nDescDistThreshold, nColorDistThreshold, nBGSamples, nRequiredBGSamples, nLBSPThresholdOffset, fRelLBSPThreshold), device, seed) {}
private:
    static lvb_params make(size_t d, size_t c, size_t n, size_t r, size_t off, float rel) {
        lvb_params p = defaults(LVB_ALGO_LOBSTER);
        p.desc_dist_threshold = (int32_t)d; p.color_dist_threshold = (int32_t)c; p.n_samples = (int32_t)n; p.n_required = (int32_t)r;
        p.lbsp_threshold_offset = (int32_t)off; p.rel_lbsp_threshold = rel;
        return p;
    }
};

/// BackgroundSubtractorPAWCS_<lv::CUDA> (ctor arguments and defaults: BackgroundSubtractorPAWCS.hpp:51-55)
struct BackgroundSubtractorPAWCS : SubtractorBase {
    explicit BackgroundSubtractorPAWCS(size_t nDescDistThresholdOffset = 2, size_t nMinColorDistThreshold = 20, size_t nMaxNbWords = 50,
                                       size_t nSamplesForMovingAvgs = 100, float fRelLBSPThreshold = 0.333f, int device = 0, uint64_t seed = 0)
        : SubtractorBase(LVB_ALGO_PAWCS, make(nDescDistThresholdOffset, nMinColorDistThreshold, nMaxNbWords, nSamplesForMovingAvgs, fRelLBSPThreshold), device, seed) {}
    /// BackgroundSubtractorPAWCS::refreshModel(nBaseOccCount, fOccDecrFrac, bForceFGUpdate) (PAWCS.cpp:107-429)
    void refreshModel(size_t nBaseOccCount, float fOccDecrFrac, bool bForceFGUpdate = false) { check(lvb_pawcs_refresh_model(m_h, (uint32_t)nBaseOccCount, fOccDecrFrac, bForceFGUpdate ? 1 : 0)); }
private:
    static lvb_params make(size_t d, size_t c, size_t n, size_t a, float rel) {
        lvb_params p = defaults(LVB_ALGO_PAWCS);
        p.desc_dist_threshold = (int32_t)d; p.color_dist_threshold = (int32_t)c; p.n_samples = (int32_t)n;
        p.n_samples_for_moving_avgs = (int32_t)a; p.rel_lbsp_threshold = rel;
        return p;
    }
};

/// BackgroundSubtractorViBe_1ch / _3ch (video/include/litiv/video/BackgroundSubtractorViBe.hpp:50-103): plain cv::BackgroundSubtractor
/// in the reference (no ROI): initialize(img), apply(img, fgmask, learningRate = 16), getBackgroundImage
template<int MODEL_CHANNELS>
#ifdef LITIV_B200_WITH_OPENCV
struct BackgroundSubtractorViBe_ : public cv::BackgroundSubtractor {
#else
struct BackgroundSubtractorViBe_ {
#endif
    explicit BackgroundSubtractorViBe_(size_t nColorDistThreshold = 20, size_t nBGSamples = 20, size_t nRequiredBGSamples = 2, int device = 0, uint64_t seed = 0) {
        check(lvb_vibe_create(MODEL_CHANNELS, (int)nColorDistThreshold, (int)nBGSamples, (int)nRequiredBGSamples, device, seed, &m_h));
    }
    virtual ~BackgroundSubtractorViBe_() { if(m_h) lvb_vibe_destroy(m_h); }
    BackgroundSubtractorViBe_(const BackgroundSubtractorViBe_&) = delete;
    BackgroundSubtractorViBe_& operator=(const BackgroundSubtractorViBe_&) = delete;
    virtual void initialize(const ImageView& img) {
        if(img.empty()) throw Exception("provided image for initialization must be non-empty and continuous");
        check(lvb_vibe_initialize(m_h, img.data, img.cols, img.rows, img.channels, img.step));
        m_rows = img.rows; m_cols = img.cols;
    }
    virtual void apply(const ImageView& img, uint8_t* fgmask, double learningRate = 16.0) {
        if(img.rows != m_rows || img.cols != m_cols) throw Exception(m_rows ? "input image size mismatch with initialization size" : "algo must be initialized first");
        if(!img.isContinuous()) throw Exception("input image data must be continuous");
        check(lvb_vibe_apply(m_h, img.data, img.channels, fgmask, learningRate));
    }
    void apply(const ImageView& img, std::vector<uint8_t>& fgmask, double learningRate = 16.0) { fgmask.resize((size_t)m_rows * m_cols); apply(img, fgmask.data(), learningRate); }
    void apply_device(const uint8_t* d_img, int channels, size_t d_step, uint8_t* d_fgmask, double learningRate = 16.0) { check(lvb_vibe_apply_device(m_h, d_img, channels, d_step, d_fgmask, learningRate)); }
    void sync() { check(lvb_vibe_sync(m_h)); }
    /// out: rows*cols*MODEL_CHANNELS bytes
    void getBackgroundImage(uint8_t* out) const { check(lvb_vibe_get_background_image(m_h, out)); }
#ifdef LITIV_B200_WITH_OPENCV
    void apply(cv::InputArray image, cv::OutputArray fgmask, double learningRate = 16.0) override {
        const cv::Mat img = image.getMat();
        fgmask.create(img.size(), CV_8UC1);
        cv::Mat m = fgmask.getMat();
        apply(ImageView(img), m.data, learningRate);
    }
    void getBackgroundImage(cv::OutputArray out) const override { out.create(m_rows, m_cols, CV_8UC(MODEL_CHANNELS)); getBackgroundImage(out.getMat().data); }
#endif
    lvb_vibe_handle handle() const { return m_h; }
protected:
    lvb_vibe_handle m_h = nullptr; int m_rows = 0, m_cols = 0;
};
typedef BackgroundSubtractorViBe_<1> BackgroundSubtractorViBe_1ch;
typedef BackgroundSubtractorViBe_<3> BackgroundSubtractorViBe_3ch;

/// BackgroundSubtractorPBAS_1ch / _3ch (video/include/litiv/video/BackgroundSubtractorPBAS.hpp:88-147): initialize(img),
/// apply(img, fgmask, learningRateOverride = -1), getBackgroundImage
template<int MODEL_CHANNELS>
#ifdef LITIV_B200_WITH_OPENCV
struct BackgroundSubtractorPBAS_ : public cv::BackgroundSubtractor {
#else
struct BackgroundSubtractorPBAS_ {
#endif
    explicit BackgroundSubtractorPBAS_(size_t nInitColorDistThreshold = 30, float fInitUpdateRate = 16.0f, size_t nBGSamples = 35, size_t nRequiredBGSamples = 2,
                                       int device = 0, uint64_t seed = 0) {
        check(lvb_pbas_create(MODEL_CHANNELS, (int)nInitColorDistThreshold, fInitUpdateRate, (int)nBGSamples, (int)nRequiredBGSamples, device, seed, &m_h));
    }
    virtual ~BackgroundSubtractorPBAS_() { if(m_h) lvb_pbas_destroy(m_h); }
    BackgroundSubtractorPBAS_(const BackgroundSubtractorPBAS_&) = delete;
    BackgroundSubtractorPBAS_& operator=(const BackgroundSubtractorPBAS_&) = delete;
    virtual void initialize(const ImageView& img) {
        if(img.empty()) throw Exception("provided image for initialization must be non-empty and continuous");
        check(lvb_pbas_initialize(m_h, img.data, img.cols, img.rows, img.channels, img.step));
        m_rows = img.rows; m_cols = img.cols;
    }
    virtual void apply(const ImageView& img, uint8_t* fgmask, double learningRateOverride = -1.0) {
        if(img.rows != m_rows || img.cols != m_cols) throw Exception(m_rows ? "input image size mismatch with initialization size" : "algo must be initialized first");
        if(!img.isContinuous()) throw Exception("input image data must be continuous");
        check(lvb_pbas_apply(m_h, img.data, img.channels, fgmask, learningRateOverride));
    }
    void apply(const ImageView& img, std::vector<uint8_t>& fgmask, double learningRateOverride = -1.0) { fgmask.resize((size_t)m_rows * m_cols); apply(img, fgmask.data(), learningRateOverride); }
    void apply_device(const uint8_t* d_img, int channels, size_t d_step, uint8_t* d_fgmask, double learningRateOverride = -1.0) { check(lvb_pbas_apply_device(m_h, d_img, channels, d_step, d_fgmask, learningRateOverride)); }
    void sync() { check(lvb_pbas_sync(m_h)); }
    void getBackgroundImage(uint8_t* out) const { check(lvb_pbas_get_background_image(m_h, out)); }
#ifdef LITIV_B200_WITH_OPENCV
    void apply(cv::InputArray image, cv::OutputArray fgmask, double learningRateOverride = -1.0) override {
        const cv::Mat img = image.getMat();
        fgmask.create(img.size(), CV_8UC1);
        cv::Mat m = fgmask.getMat();
        apply(ImageView(img), m.data, learningRateOverride);
    }
    void getBackgroundImage(cv::OutputArray out) const override { out.create(m_rows, m_cols, CV_8UC(MODEL_CHANNELS)); getBackgroundImage(out.getMat().data); }
#endif
    lvb_pbas_handle handle() const { return m_h; }
protected:
    lvb_pbas_handle m_h = nullptr; int m_rows = 0, m_cols = 0;
};
typedef BackgroundSubtractorPBAS_<1> BackgroundSubtractorPBAS_1ch;
typedef BackgroundSubtractorPBAS_<3> BackgroundSubtractorPBAS_3ch;

/// lv::BinClassif (datasets/include/litiv/datasets/metrics.hpp:32-67) with accumulate() on the device, and BinClassifMetrics (:213-257)
struct BinClassif {
    uint64_t nTP = 0, nTN = 0, nFP = 0, nFN = 0, nSE = 0, nDC = 0;
    uint64_t total(bool bWithDontCare = false) const { return nTP + nTN + nFP + nFN + (bWithDontCare ? nDC : uint64_t(0)); }
    /// scores the subtractor's latest foreground mask where it lives (no mask read-back)
    void accumulate(const SubtractorBase& algo, const ImageView& gt, const ImageView& roi = ImageView()) {
        uint64_t c[6] = {nTP, nTN, nFP, nFN, nSE, nDC};
        check(lvb_binclassif_accumulate(algo.handle(), gt.empty() ? nullptr : gt.data, roi.empty() ? nullptr : roi.data, c));
        nTP = c[0]; nTN = c[1]; nFP = c[2]; nFN = c[3]; nSE = c[4]; nDC = c[5];
    }
    void accumulate(const ImageView& classif, const ImageView& gt, const ImageView& roi = ImageView(), int device = 0) {
        if(classif.empty() || classif.channels != 1 || !classif.isContinuous()) throw Exception("binary classifier results must be non-empty and of type 8UC1");
        if((!gt.empty() && (gt.rows != classif.rows || gt.cols != classif.cols)) || (!roi.empty() && (roi.rows != classif.rows || roi.cols != classif.cols)))
            throw Exception("all input mat sizes must match");
        uint64_t c[6] = {nTP, nTN, nFP, nFN, nSE, nDC};
        check(lvb_binclassif(classif.data, gt.empty() ? nullptr : gt.data, roi.empty() ? nullptr : roi.data, classif.cols, classif.rows, c, device));
        nTP = c[0]; nTN = c[1]; nFP = c[2]; nFN = c[3]; nSE = c[4]; nDC = c[5];
    }
};
struct BinClassifMetrics {
    double dRecall, dSpecificity, dFPR, dFNR, dPBC, dPrecision, dFMeasure, dMCC;
    explicit BinClassifMetrics(const BinClassif& m) {
        const uint64_t c[6] = {m.nTP, m.nTN, m.nFP, m.nFN, m.nSE, m.nDC};
        double o[8];
        check(lvb_binclassif_metrics(c, o));
        dRecall = o[0]; dSpecificity = o[1]; dFPR = o[2]; dFNR = o[3]; dPBC = o[4]; dPrecision = o[5]; dFMeasure = o[6]; dMCC = o[7];
    }
};

/// LBSP dense extractor (features2d LBSP: LBSP(size_t) absolute / LBSP(float, size_t) relative; compute2; setReference)
struct LBSP {
    explicit LBSP(size_t nThreshold, int device = 0) : m_abs(true), m_rel(0), m_thr((int)nThreshold), m_dev(device) {}
    LBSP(float fRelThreshold, size_t nThresholdOffset, int device = 0) : m_abs(false), m_rel(fRelThreshold), m_thr((int)nThresholdOffset), m_dev(device) {
        if(fRelThreshold < 0) throw Exception("relative LBSP threshold must be non-negative");
    }
    static constexpr int PATCH_SIZE = 5, DESC_SIZE = 2, DESC_SIZE_BITS = 16;
    /// only two dimensions exist (features2d/test/lbsp.cpp:13-14: borderSize(2) throws)
    int borderSize(int nDim = 0) const { if(nDim < 0 || nDim > 1) throw Exception("border size is only defined for 2 dimensions"); return PATCH_SIZE / 2; }
    int windowSize() const { return PATCH_SIZE; }   // square window: width == height == 5
    int descriptorSize() const { return DESC_SIZE; }
    int descriptorType() const { return 2; }         // CV_16U
    int defaultNorm() const { return 6; }            // cv::NORM_HAMMING
    void setReference(const ImageView& ref) { m_ref = ref; }
    /// out: rows*cols*channels uint16; the 2-px border is left untouched like the reference's oDesc.create()
    void compute2(const ImageView& img, uint16_t* out) const {
        if(!m_ref.empty() && (m_ref.rows != img.rows || m_ref.cols != img.cols || m_ref.channels != img.channels))
            throw Exception("ref image must be empty, or of the same size/type as the input image");
        check(lvb_lbsp_compute(img.data, m_ref.empty() ? nullptr : m_ref.data, img.cols, img.rows, img.channels, m_abs ? 0 : 1, m_rel, m_thr, out, m_dev));
    }
private:
    bool m_abs; float m_rel; int m_thr, m_dev; ImageView m_ref;
};

/// EdgeDetectorLBSP (imgproc/include/litiv/imgproc/EdgeDetectorLBSP.hpp:33-83): same constructor defaults and method names; masks are
/// caller-owned rows*cols bytes. Like the reference object, the detector keeps its maps between calls and is not thread-safe.
struct EdgeDetectorLBSP {
    /// bNormalizeOutput (EdgeDetectorLBSP.cpp:431-432): apply() min-max normalises its confidence map like cv::normalize(NORM_MINMAX)
    explicit EdgeDetectorLBSP(size_t nLevels = 3, double dHystLowThrshFactor = 0.5, bool bNormalizeOutput = false, int device = 0) {
        check(lvb_edge_create((int)nLevels, dHystLowThrshFactor, device, &m_h));
        if(bNormalizeOutput) check(lvb_edge_set_normalize(m_h, 1));
    }
    ~EdgeDetectorLBSP() { lvb_edge_destroy(m_h); }
    EdgeDetectorLBSP(const EdgeDetectorLBSP&) = delete;
    EdgeDetectorLBSP& operator=(const EdgeDetectorLBSP&) = delete;
    double getDefaultThreshold() const { return lvb_edge_default_threshold(); }
    void apply_threshold(const ImageView& img, uint8_t* oEdgeMask, double dDetThreshold = 8.0 / 16.0) {
        checkInput(img);
        check(lvb_edge_apply_threshold(m_h, img.data, img.cols, img.rows, img.channels, oEdgeMask, dDetThreshold));
    }
    void apply(const ImageView& img, uint8_t* oEdgeMask) {
        checkInput(img);
        check(lvb_edge_apply(m_h, img.data, img.cols, img.rows, img.channels, oEdgeMask));
    }
#ifdef LITIV_B200_WITH_OPENCV
    void apply_threshold(cv::InputArray oInputImage, cv::OutputArray oEdgeMask, double dDetThreshold = 8.0 / 16.0) {
        const cv::Mat m = oInputImage.getMat();
        oEdgeMask.create(m.size(), CV_8UC1);
        apply_threshold(ImageView(m), oEdgeMask.getMat().data, dDetThreshold);
    }
    void apply(cv::InputArray oInputImage, cv::OutputArray oEdgeMask) {
        const cv::Mat m = oInputImage.getMat();
        oEdgeMask.create(m.size(), CV_8UC1);
        apply(ImageView(m), oEdgeMask.getMat().data);
    }
#endif
    lvb_edge_handle handle() const { return m_h; }
private:
    static void checkInput(const ImageView& img) {
        if(img.empty() || !img.isContinuous() || img.channels < 1 || img.channels > 4) throw Exception("input image must be non-empty and continuous, 8UC1 .. 8UC4");
    }
    lvb_edge_handle m_h = nullptr;
};

} // namespace lvb
