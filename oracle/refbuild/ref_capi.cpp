// ORACLE — TEST INFRASTRUCTURE ONLY.
// C entry points over the REFERENCE'S OWN classes (BackgroundSubtractorSuBSENSE / LOBSTER / PAWCS, LBSP), compiled unmodified from
// /root/reference against the cvcompat headers (oracle/cvcompat). Built into oracle/_ref/liblitiv_ref.so by `make -C oracle _ref`.
// Purpose: pin the oracle's reference-order mode (oracle/lvo_*.hpp, MODE_REFERENCE, glibc rand() clone) bit-for-bit to the reference
// source itself (tests/test_ref_pin_cpu.py), and serve as the CPU baseline of bench.py (`cpu_baseline.kind = "reference"`).
// The reference draws from the process-global rand(): ref_create() calls srand(seed), so drive ONE instance at a time.
#include "litiv/video/BackgroundSubtractorSuBSENSE.hpp"
#include "litiv/video/BackgroundSubtractorLOBSTER.hpp"
#include "litiv/video/BackgroundSubtractorPAWCS.hpp"
#include "litiv/features2d/LBSP.hpp"
#include "litiv/video/BackgroundSubtractorViBe.hpp"
#include "litiv/video/BackgroundSubtractorPBAS.hpp"
#include "litiv/imgproc/EdgeDetectorLBSP.hpp"
#include <chrono>
#include <cstring>
#include <string>
#include <vector>

namespace {
thread_local std::string g_err;

struct RefParams { // same layout as lvo::Params / lvb_params
    float rel_lbsp_threshold; int lbsp_threshold_offset, desc_dist_threshold, color_dist_threshold, n_samples, n_required, n_samples_for_moving_avgs,
        n_global_words, median_blur_kernel_size;
};
typedef std::vector<unsigned char> Bytes;

template<typename T> void put_mat(const cv::Mat& m, Bytes& out) {
    const size_t rowbytes = (size_t)m.cols * m.elemSize();
    out.resize(rowbytes * m.rows);
    for(int y = 0; y < m.rows; ++y) std::memcpy(out.data() + rowbytes * y, m.ptr(y), rowbytes);
}
void put_mat(const cv::Mat& m, Bytes& out) { put_mat<unsigned char>(m, out); }
void put_samples(const std::vector<cv::Mat>& v, Bytes& out) { // [N][H*W*C]
    out.clear();
    for(const cv::Mat& m : v) { Bytes b; put_mat(m, b); out.insert(out.end(), b.begin(), b.end()); }
}
template<typename T> void put_vec(const std::vector<T>& v, Bytes& out) { out.resize(v.size() * sizeof(T)); if(!v.empty()) std::memcpy(out.data(), v.data(), out.size()); }

struct Base {
    virtual ~Base() {}
    virtual IBackgroundSubtractorLBSP& algo() = 0;
    virtual bool get(const std::string& n, Bytes& out) = 0;
    virtual void scalars(double* d) = 0;
    int W = 0, H = 0, C = 0;
};

struct RefSub : BackgroundSubtractorSuBSENSE, Base {
    using BackgroundSubtractorSuBSENSE::BackgroundSubtractorSuBSENSE;
    IBackgroundSubtractorLBSP& algo() override { return *this; }
    bool get(const std::string& n, Bytes& out) override {
#define M(name, mat) if(n == name) { put_mat(mat, out); return true; }
        M("roi", m_oROI) M("lastfg", m_oLastFGMask) M("lastcolor", m_oLastColorFrame) M("lastdesc", m_oLastDescFrame)
        M("T", m_oUpdateRateFrame) M("R", m_oDistThresholdFrame) M("v", m_oVariationModulatorFrame) M("Dlast", m_oMeanLastDistFrame)
        M("DminLT", m_oMeanMinDistFrame_LT) M("DminST", m_oMeanMinDistFrame_ST) M("rawLT", m_oMeanRawSegmResFrame_LT) M("rawST", m_oMeanRawSegmResFrame_ST)
        M("finLT", m_oMeanFinalSegmResFrame_LT) M("finST", m_oMeanFinalSegmResFrame_ST) M("dsLT", m_oMeanDownSampledLastDistFrame_LT)
        M("dsST", m_oMeanDownSampledLastDistFrame_ST) M("unstable", m_oUnstableRegionMask) M("blinks", m_oBlinksFrame) M("lastraw", m_oLastRawFGMask)
        M("lastrawblink", m_oLastRawFGBlinkMask) M("dilinv", m_oLastFGMask_dilated_inverted)
#undef M
        if(n == "lut") { out.assign(m_anLBSPThreshold_8bitLUT.begin(), m_anLBSPThreshold_8bitLUT.end()); return true; }
        if(n == "bg_color") { put_samples(m_voBGColorSamples, out); return true; }
        if(n == "bg_desc") { put_samples(m_voBGDescSamples, out); return true; }
        return false;
    }
    void scalars(double* d) override {
        d[0] = (double)m_nFrameIdx; d[1] = (double)m_nFramesSinceLastReset; d[2] = (double)m_nModelResetCooldown; d[3] = m_bAutoModelResetEnabled;
        d[4] = m_bLearningRateScalingEnabled; d[5] = m_bUse3x3Spread; d[6] = m_nMedianBlurKernelSize; d[7] = m_fCurrLearningRateLowerCap;
        d[8] = m_fCurrLearningRateUpperCap; d[9] = m_fLastNonZeroDescRatio; d[10] = (double)m_nFinalROIPxCount; d[11] = (double)m_nOrigROIPxCount;
    }
};

struct RefLob : BackgroundSubtractorLOBSTER, Base {
    using BackgroundSubtractorLOBSTER::BackgroundSubtractorLOBSTER;
    IBackgroundSubtractorLBSP& algo() override { return *this; }
    bool get(const std::string& n, Bytes& out) override {
        if(n == "roi") { put_mat(m_oROI, out); return true; }
        if(n == "lastfg") { put_mat(m_oLastFGMask, out); return true; }
        if(n == "lastcolor") { put_mat(m_oLastColorFrame, out); return true; }
        if(n == "lastdesc") { put_mat(m_oLastDescFrame, out); return true; }
        if(n == "lut") { out.assign(m_anLBSPThreshold_8bitLUT.begin(), m_anLBSPThreshold_8bitLUT.end()); return true; }
        if(n == "bg_color") { put_samples(m_voBGColorSamples, out); return true; }
        if(n == "bg_desc") { put_samples(m_voBGDescSamples, out); return true; }
        return false;
    }
    void scalars(double* d) override { d[0] = (double)m_nFrameIdx; d[3] = m_bAutoModelResetEnabled; d[10] = (double)m_nFinalROIPxCount; d[11] = (double)m_nOrigROIPxCount; }
};

struct RefPaw : BackgroundSubtractorPAWCS, Base {
    using BackgroundSubtractorPAWCS::BackgroundSubtractorPAWCS;
    IBackgroundSubtractorLBSP& algo() override { return *this; }
    size_t gid(const GlobalWordBase* p) const { // identity of a global word = its position in the word list (creation order)
        if(m_nImgChannels == 1) return (size_t)((const GlobalWord_1ch*)p - m_voGlobalWordList_1ch.data());
        return (size_t)((const GlobalWord_3ch*)p - m_voGlobalWordList_3ch.data());
    }
    bool get(const std::string& n, Bytes& out) override {
#define M(name, mat) if(n == name) { put_mat(mat, out); return true; }
        M("roi", m_oROI) M("lastfg", m_oLastFGMask) M("lastcolor", m_oLastColorFrame) M("lastdesc", m_oLastDescFrame)
        M("T", m_oUpdateRateFrame) M("R", m_oDistThresholdFrame) M("v", m_oDistThresholdVariationFrame)
        M("DminLT", m_oMeanMinDistFrame_LT) M("DminST", m_oMeanMinDistFrame_ST) M("rawLT", m_oMeanRawSegmResFrame_LT) M("rawST", m_oMeanRawSegmResFrame_ST)
        M("finLT", m_oMeanFinalSegmResFrame_LT) M("finST", m_oMeanFinalSegmResFrame_ST) M("dsLT", m_oMeanDownSampledLastDistFrame_LT)
        M("dsST", m_oMeanDownSampledLastDistFrame_ST) M("unstable", m_oUnstableRegionMask) M("illum", m_oIllumUpdtRegionMask) M("blinks", m_oBlinksFrame)
        M("lastraw", m_oLastRawFGMask) M("lastrawblink", m_oLastRawFGBlinkMask) M("dil", m_oLastFGMask_dilated) M("dilinv", m_oLastFGMask_dilated_inverted)
#undef M
        if(n == "lut") { out.assign(m_anLBSPThreshold_8bitLUT.begin(), m_anLBSPThreshold_8bitLUT.end()); return true; }
        const size_t npx = m_nTotPxCount, NW = m_nCurrLocalWords, NG = m_nCurrGlobalWords, Cn = m_nImgChannels;
        const bool lw = n == "lw_first" || n == "lw_last" || n == "lw_occ" || n == "lw_color" || n == "lw_desc" || n == "lw_valid";
        if(lw) { // [p*NW + i] in dictionary order, zeros where the pixel is outside the ROI or the slot is empty
            std::vector<uint32_t> a(npx * NW, 0); std::vector<unsigned char> col(npx * NW * Cn, 0), valid(npx * NW, 0); std::vector<unsigned short> des(npx * NW * Cn, 0);
            for(size_t p = 0; p < npx; ++p) {
                const size_t mi = m_voPxInfoLUT_PAWCS[p].nModelIdx;
                if(!m_oROI.data[p]) continue;
                for(size_t i = 0; i < NW; ++i) {
                    const LocalWordBase* w = m_vpLocalWordDict[mi * NW + i];
                    if(!w) continue;
                    const size_t k = p * NW + i;
                    valid[k] = 1;
                    a[k] = (uint32_t)(n == "lw_first" ? w->nFirstOcc : n == "lw_last" ? w->nLastOcc : w->nOccurrences);
                    for(size_t c = 0; c < Cn; ++c) {
                        if(Cn == 1) { col[k] = ((const LocalWord_1ch*)w)->oFeature.anColor[0]; des[k] = ((const LocalWord_1ch*)w)->oFeature.anDesc[0]; }
                        else { col[k * 3 + c] = ((const LocalWord_3ch*)w)->oFeature.anColor[c]; des[k * 3 + c] = ((const LocalWord_3ch*)w)->oFeature.anDesc[c]; }
                    }
                }
            }
            if(n == "lw_color") put_vec(col, out); else if(n == "lw_desc") put_vec(des, out); else if(n == "lw_valid") put_vec(valid, out); else put_vec(a, out);
            return true;
        }
        if(n == "gdict") { std::vector<int32_t> g(NG, -1); for(size_t i = 0; i < NG; ++i) if(m_vpGlobalWordDict[i]) g[i] = (int32_t)gid(m_vpGlobalWordDict[i]); put_vec(g, out); return true; }
        if(n == "gw_weight" || n == "gw_bits" || n == "gw_color" || n == "gw_desc" || n == "gw_map") { // indexed by DICTIONARY position (identity order is an oracle-internal choice)
            const size_t gpx = (size_t)m_oDownSampledFrameSize_GlobalWordLookup.area();
            std::vector<float> wgt(NG, 0.f), map(NG * gpx, 0.f); std::vector<unsigned char> bits(NG, 0), col(NG * Cn, 0); std::vector<unsigned short> des(NG * Cn, 0);
            for(size_t i = 0; i < NG; ++i) {
                const GlobalWordBase* w = m_vpGlobalWordDict[i];
                if(!w) continue;
                wgt[i] = w->fLatestWeight; bits[i] = w->nDescBITS;
                if(!w->oSpatioOccMap.empty()) for(size_t k = 0; k < gpx; ++k) map[i * gpx + k] = ((const float*)w->oSpatioOccMap.data)[k];
                for(size_t c = 0; c < Cn; ++c) {
                    if(Cn == 1) { col[i] = ((const GlobalWord_1ch*)w)->oFeature.anColor[0]; des[i] = ((const GlobalWord_1ch*)w)->oFeature.anDesc[0]; }
                    else { col[i * 3 + c] = ((const GlobalWord_3ch*)w)->oFeature.anColor[c]; des[i * 3 + c] = ((const GlobalWord_3ch*)w)->oFeature.anDesc[c]; }
                }
            }
            if(n == "gw_weight") put_vec(wgt, out); else if(n == "gw_bits") put_vec(bits, out); else if(n == "gw_color") put_vec(col, out); else if(n == "gw_desc") put_vec(des, out); else put_vec(map, out);
            return true;
        }
        if(n == "glut") { // per pixel: dictionary POSITION of the i-th entry of its sort LUT (0xFF outside the ROI)
            std::vector<unsigned char> g(npx * NG, 0xFF);
            for(size_t p = 0; p < npx; ++p) {
                if(!m_oROI.data[p]) continue;
                const auto& lut = m_voPxInfoLUT_PAWCS[p].vpGlobalDictSortLUT;
                for(size_t i = 0; i < NG && i < lut.size(); ++i)
                    for(size_t j = 0; j < NG; ++j) if(m_vpGlobalWordDict[j] == lut[i]) { g[p * NG + i] = (unsigned char)j; break; }
            }
            put_vec(g, out); return true;
        }
        return false;
    }
    void scalars(double* d) override {
        d[0] = (double)m_nFrameIdx; d[1] = (double)m_nFramesSinceLastReset; d[2] = (double)m_nModelResetCooldown; d[3] = m_bAutoModelResetEnabled;
        d[4] = (double)m_nCurrLocalWords; d[5] = (double)m_nCurrGlobalWords; d[6] = m_nMedianBlurKernelSize; d[7] = (double)m_nLocalWordWeightOffset;
        d[8] = m_bUsingMovingCamera; d[9] = m_fLastNonFlatRegionRatio; d[10] = (double)m_nFinalROIPxCount; d[11] = (double)m_nOrigROIPxCount;
    }
};

struct Handle { int algo; Base* b; };
} // namespace

#define REF_TRY try {
#define REF_CATCH } catch(const std::exception& e) { g_err = e.what(); return 1; } catch(...) { g_err = "unknown exception"; return 1; } return 0;

extern "C" {

const char* ref_last_error() { return g_err.c_str(); }
const char* ref_version() { return "litiv reference sources, compiled unmodified against oracle/cvcompat (LITIV_VERSION " LITIV_VERSION_STR ")"; }

/// algo: 0 LOBSTER, 1 SuBSENSE, 2 PAWCS (ids of include/litiv_b200.h). params NULL = the reference's defaults. Seeds the global rand().
int ref_create(int algo, const RefParams* p, unsigned seed, void** out) {
    REF_TRY
    Handle* h = new Handle{algo, nullptr};
    if(algo == 1) h->b = p ? new RefSub((size_t)p->desc_dist_threshold, (size_t)p->color_dist_threshold, (size_t)p->n_samples, (size_t)p->n_required, (size_t)p->n_samples_for_moving_avgs, p->rel_lbsp_threshold) : new RefSub();
    else if(algo == 0) h->b = p ? new RefLob((size_t)p->desc_dist_threshold, (size_t)p->color_dist_threshold, (size_t)p->n_samples, (size_t)p->n_required, (size_t)p->lbsp_threshold_offset, p->rel_lbsp_threshold) : new RefLob();
    else if(algo == 2) h->b = p ? new RefPaw((size_t)p->desc_dist_threshold, (size_t)p->color_dist_threshold, (size_t)p->n_samples, (size_t)p->n_samples_for_moving_avgs, p->rel_lbsp_threshold) : new RefPaw();
    else { delete h; throw std::runtime_error("unknown algorithm id"); }
    srand(seed);
    *out = h;
    REF_CATCH
}
int ref_destroy(void* hv) { Handle* h = (Handle*)hv; if(h) { delete h->b; delete h; } return 0; }

int ref_initialize(void* hv, const unsigned char* img, int w, int h, int c, const unsigned char* roi) {
    REF_TRY
    Handle* H = (Handle*)hv;
    cv::Mat im(h, w, CV_8UC(c), (void*)img), r;
    if(roi) r = cv::Mat(h, w, CV_8UC1, (void*)roi);
    H->b->W = w; H->b->H = h; H->b->C = c;
    H->b->algo().initialize(im.clone(), r.empty() ? cv::Mat() : r.clone());
    REF_CATCH
}
int ref_apply(void* hv, const unsigned char* img, unsigned char* mask, double lr) {
    REF_TRY
    Handle* H = (Handle*)hv; Base* b = H->b;
    cv::Mat im(b->H, b->W, CV_8UC(b->C), (void*)img), m(b->H, b->W, CV_8UC1, mask);
    b->algo().apply(im, m, lr);
    if(m.data != mask) throw std::runtime_error("apply() reallocated the output mask");
    REF_CATCH
}
/// n frames back to back; returns the seconds spent inside apply() (negative on error). lrs[n].
double ref_apply_sequence(void* hv, const unsigned char* frames, int n, size_t frame_bytes, unsigned char* last_mask, const double* lrs) {
    try {
        Handle* H = (Handle*)hv; Base* b = H->b;
        cv::Mat m(b->H, b->W, CV_8UC1, last_mask);
        double t = 0;
        for(int i = 0; i < n; ++i) {
            cv::Mat im(b->H, b->W, CV_8UC(b->C), (void*)(frames + frame_bytes * i));
            const auto t0 = std::chrono::steady_clock::now();
            b->algo().apply(im, m, lrs[i]);
            t += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        }
        return t;
    } catch(const std::exception& e) { g_err = e.what(); return -1.0; }
}
int ref_refresh_model(void* hv, float frac, int force_fg) {
    REF_TRY
    Handle* H = (Handle*)hv;
    if(H->algo == 1) static_cast<RefSub*>(H->b)->refreshModel(frac, force_fg != 0);
    else if(H->algo == 0) static_cast<RefLob*>(H->b)->refreshModel(frac, force_fg != 0);
    else throw std::runtime_error("use ref_pawcs_refresh_model");
    REF_CATCH
}
int ref_pawcs_refresh_model(void* hv, unsigned long long base_occ, float decr_frac, int force_fg) {
    REF_TRY
    Handle* H = (Handle*)hv;
    if(H->algo != 2) throw std::runtime_error("not a PAWCS instance");
    static_cast<RefPaw*>(H->b)->refreshModel((size_t)base_occ, decr_frac, force_fg != 0);
    REF_CATCH
}
int ref_set_auto_model_reset(void* hv, int v) { REF_TRY ((Handle*)hv)->b->algo().setAutomaticModelReset(v != 0); REF_CATCH }
int ref_get_background_image(void* hv, unsigned char* out) {
    REF_TRY
    Base* b = ((Handle*)hv)->b; cv::Mat m;
    b->algo().getBackgroundImage(m);
    if(m.type() != CV_8UC(b->C) || m.rows != b->H || m.cols != b->W) throw std::runtime_error("unexpected background image type");
    Bytes t; put_mat(m, t); std::memcpy(out, t.data(), t.size());
    REF_CATCH
}
int ref_get_background_descriptors_image(void* hv, unsigned short* out) {
    REF_TRY
    Base* b = ((Handle*)hv)->b; cv::Mat m;
    b->algo().getBackgroundDescriptorsImage(m);
    if(m.type() != CV_16UC(b->C) || m.rows != b->H || m.cols != b->W) throw std::runtime_error("unexpected background descriptor image type");
    Bytes t; put_mat(m, t); std::memcpy(out, t.data(), t.size());
    REF_CATCH
}
int ref_get_roi(void* hv, unsigned char* out) {
    REF_TRY
    Base* b = ((Handle*)hv)->b; cv::Mat m = b->algo().getROICopy();
    Bytes t; put_mat(m, t); std::memcpy(out, t.data(), t.size());
    REF_CATCH
}
/// state buffers by the oracle's names; two calls: out == NULL returns the size in *bytes
int ref_state_get(void* hv, const char* name, void* out, size_t* bytes) {
    REF_TRY
    Base* b = ((Handle*)hv)->b;
    if(std::string(name) == "scalars") {
        if(!out) { *bytes = 16 * sizeof(double); return 0; }
        double d[16] = {0}; b->scalars(d); std::memcpy(out, d, sizeof(d)); return 0;
    }
    Bytes t;
    if(!b->get(name, t)) throw std::runtime_error(std::string("unknown state buffer: ") + name);
    if(!out) { *bytes = t.size(); return 0; }
    if(*bytes != t.size()) throw std::runtime_error(std::string("size mismatch for state buffer ") + name);
    std::memcpy(out, t.data(), t.size());
    REF_CATCH
}

/// LBSP::compute2 (dense) through the reference's extractor: rel < 0 -> absolute threshold `thr`, else relative `rel` with offset `thr`
int ref_lbsp_compute(const unsigned char* img, const unsigned char* ref_or_null, int w, int h, int c, int use_rel, float rel, int thr, unsigned short* out) {
    REF_TRY
    cv::Mat im(h, w, CV_8UC(c), (void*)img), d;
    std::unique_ptr<LBSP> e(use_rel ? new LBSP(rel, (size_t)thr) : new LBSP((size_t)thr));
    if(ref_or_null) e->setReference(cv::Mat(h, w, CV_8UC(c), (void*)ref_or_null));
    e->compute2(im, d);
    if(d.type() != CV_16UC(c) || d.rows != h || d.cols != w) throw std::runtime_error("unexpected descriptor map type");
    // the reference leaves the 2-px border uninitialised (oDesc.create, never written): zero it so that callers can compare whole maps
    for(int y = 0; y < h; ++y) for(int x = 0; x < w; ++x) for(int k = 0; k < c; ++k) {
        const bool border = x < 2 || y < 2 || x >= w - 2 || y >= h - 2;
        out[((size_t)y * w + x) * c + k] = border ? 0 : d.ptr<unsigned short>(y)[x * c + k];
    }
    REF_CATCH
}

/// helper known answers straight from the reference's headers (utils/math.hpp, utils/opencv.hpp)
unsigned long long ref_cdist3(const unsigned char* a, const unsigned char* b) { return (unsigned long long)lv::cdist<3>(a, b); }
unsigned ref_L1dist3_u8(const unsigned char* a, const unsigned char* b) { return (unsigned)lv::L1dist<3>(a, b); }
unsigned ref_hdist3(const unsigned short* a, const unsigned short* b) { return (unsigned)lv::hdist<3>(a, b); }
void ref_sample_pos_7x7(int rnd, int ox, int oy, int border, int w, int h, int* sx, int* sy) { lv::getSamplePosition_7x7_std2(rnd, *sx, *sy, ox, oy, border, cv::Size(w, h)); }
void ref_neighbor_pos_3x3(int rnd, int ox, int oy, int border, int w, int h, int* nx, int* ny) { lv::getNeighborPosition_3x3(rnd, *nx, *ny, ox, oy, border, cv::Size(w, h)); }
void ref_neighbor_pos_5x5(int rnd, int ox, int oy, int border, int w, int h, int* nx, int* ny) { lv::getNeighborPosition_5x5(rnd, *nx, *ny, ox, oy, border, cv::Size(w, h)); }

} // extern "C"

// ---- ViBe / PBAS: the reference's own BackgroundSubtractorViBe_1ch/_3ch and BackgroundSubtractorPBAS_1ch/_3ch ----
namespace {
struct VibeBase { virtual ~VibeBase() {} virtual BackgroundSubtractorViBe& algo() = 0; virtual const std::vector<cv::Mat>& model() const = 0; int W = 0, H = 0, C = 0; };
struct RefViBe1 : BackgroundSubtractorViBe_1ch, VibeBase { using BackgroundSubtractorViBe_1ch::BackgroundSubtractorViBe_1ch;
    BackgroundSubtractorViBe& algo() override { return *this; } const std::vector<cv::Mat>& model() const override { return m_voBGImg; } };
struct RefViBe3 : BackgroundSubtractorViBe_3ch, VibeBase { using BackgroundSubtractorViBe_3ch::BackgroundSubtractorViBe_3ch;
    BackgroundSubtractorViBe& algo() override { return *this; } const std::vector<cv::Mat>& model() const override { return m_voBGImg; } };
struct PbasBase { virtual ~PbasBase() {} virtual BackgroundSubtractorPBAS& algo() = 0; virtual bool get(const std::string& n, Bytes& out) = 0; int W = 0, H = 0, C = 0; };
#define PBAS_GET \
    bool get(const std::string& n, Bytes& out) override { \
        if(n == "bg_color") { put_samples(m_voBGImg, out); return true; } \
        if(n == "bg_grad") { put_samples(m_voBGGrad, out); return true; } \
        if(n == "R") { put_mat(m_oDistThresholdFrame, out); return true; } \
        if(n == "T") { put_mat(m_oUpdateRateFrame, out); return true; } \
        if(n == "meanmin") { put_mat(m_oMeanMinDistFrame, out); return true; } \
        if(n == "scalars") { const double d[2] = {0.0, (double)m_fFormerMeanGradDist}; out.assign((const unsigned char*)d, (const unsigned char*)d + sizeof(d)); return true; } \
        return false; }
struct RefPBAS1 : BackgroundSubtractorPBAS_1ch, PbasBase { using BackgroundSubtractorPBAS_1ch::BackgroundSubtractorPBAS_1ch;
    BackgroundSubtractorPBAS& algo() override { return *this; } PBAS_GET };
struct RefPBAS3 : BackgroundSubtractorPBAS_3ch, PbasBase { using BackgroundSubtractorPBAS_3ch::BackgroundSubtractorPBAS_3ch;
    BackgroundSubtractorPBAS& algo() override { return *this; } PBAS_GET };
}

extern "C" {
int ref_vibe_create(int model_channels, int color_dist_threshold, int n_samples, int n_required, unsigned seed, void** out) {
    REF_TRY
    if(model_channels != 1 && model_channels != 3) throw std::runtime_error("model channels must be 1 or 3");
    srand(seed);
    VibeBase* b = model_channels == 1 ? (VibeBase*)new RefViBe1((size_t)color_dist_threshold, (size_t)n_samples, (size_t)n_required)
                                      : (VibeBase*)new RefViBe3((size_t)color_dist_threshold, (size_t)n_samples, (size_t)n_required);
    b->C = model_channels; *out = b;
    REF_CATCH
}
int ref_vibe_destroy(void* h) { delete (VibeBase*)h; return 0; }
int ref_vibe_initialize(void* h, const unsigned char* img, int w, int hh, int c) {
    REF_TRY
    VibeBase* b = (VibeBase*)h; b->W = w; b->H = hh;
    b->algo().initialize(cv::Mat(hh, w, CV_8UC(c), (void*)img).clone());
    REF_CATCH
}
int ref_vibe_apply(void* h, const unsigned char* img, int c, unsigned char* mask, double lr) {
    REF_TRY
    VibeBase* b = (VibeBase*)h; cv::Mat m;
    b->algo().apply(cv::Mat(b->H, b->W, CV_8UC(c), (void*)img), m, lr);
    if(m.type() != CV_8UC1 || m.rows != b->H || m.cols != b->W) throw std::runtime_error("unexpected mask type");
    Bytes t; put_mat(m, t); std::memcpy(mask, t.data(), t.size());
    REF_CATCH
}
int ref_vibe_model(void* h, unsigned char* out, size_t bytes) {
    REF_TRY
    Bytes t; put_samples(((VibeBase*)h)->model(), t);
    if(t.size() != bytes) throw std::runtime_error("size mismatch for the ViBe model");
    std::memcpy(out, t.data(), bytes);
    REF_CATCH
}
int ref_vibe_get_background_image(void* h, unsigned char* out) {
    REF_TRY
    VibeBase* b = (VibeBase*)h; cv::Mat m; b->algo().getBackgroundImage(m);
    if(m.type() != CV_8UC(b->C) || m.rows != b->H || m.cols != b->W) throw std::runtime_error("unexpected background image type");
    Bytes t; put_mat(m, t); std::memcpy(out, t.data(), t.size());
    REF_CATCH
}
int ref_pbas_create(int model_channels, int color_dist_threshold, float update_rate, int n_samples, int n_required, unsigned seed, void** out) {
    REF_TRY
    if(model_channels != 1 && model_channels != 3) throw std::runtime_error("model channels must be 1 or 3");
    srand(seed);
    PbasBase* b = model_channels == 1 ? (PbasBase*)new RefPBAS1((size_t)color_dist_threshold, update_rate, (size_t)n_samples, (size_t)n_required)
                                      : (PbasBase*)new RefPBAS3((size_t)color_dist_threshold, update_rate, (size_t)n_samples, (size_t)n_required);
    b->C = model_channels; *out = b;
    REF_CATCH
}
int ref_pbas_destroy(void* h) { delete (PbasBase*)h; return 0; }
int ref_pbas_initialize(void* h, const unsigned char* img, int w, int hh, int c) {
    REF_TRY
    PbasBase* b = (PbasBase*)h; b->W = w; b->H = hh;
    b->algo().initialize(cv::Mat(hh, w, CV_8UC(c), (void*)img).clone());
    REF_CATCH
}
int ref_pbas_apply(void* h, const unsigned char* img, int c, unsigned char* mask, double lr) {
    REF_TRY
    PbasBase* b = (PbasBase*)h; cv::Mat m;
    b->algo().apply(cv::Mat(b->H, b->W, CV_8UC(c), (void*)img), m, lr);
    if(m.type() != CV_8UC1 || m.rows != b->H || m.cols != b->W) throw std::runtime_error("unexpected mask type");
    Bytes t; put_mat(m, t); std::memcpy(mask, t.data(), t.size());
    REF_CATCH
}
int ref_pbas_state_get(void* h, const char* name, void* out, size_t bytes) {
    REF_TRY
    Bytes t;
    if(!((PbasBase*)h)->get(name, t)) throw std::runtime_error(std::string("unknown state buffer: ") + name);
    if(t.size() != bytes) throw std::runtime_error(std::string("size mismatch for state buffer ") + name);
    std::memcpy(out, t.data(), bytes);
    REF_CATCH
}
int ref_pbas_get_background_image(void* h, unsigned char* out) {
    REF_TRY
    PbasBase* b = (PbasBase*)h; cv::Mat m; b->algo().getBackgroundImage(m);
    if(m.type() != CV_8UC(b->C) || m.rows != b->H || m.cols != b->W) throw std::runtime_error("unexpected background image type");
    Bytes t; put_mat(m, t); std::memcpy(out, t.data(), t.size());
    REF_CATCH
}
} // extern "C"

// ---- EdgeDetectorLBSP: the reference's own imgproc/src/EdgeDetectorLBSP.cpp ----
namespace {
struct RefEdge : EdgeDetectorLBSP {
    using EdgeDetectorLBSP::EdgeDetectorLBSP;
    const lv::aligned_vector<uchar,32>& grad() const { return m_vuLBSPGradMapData; }
    const lv::aligned_vector<uchar,32>& edge() const { return m_vuEdgeTempMaskData; }
};
}
extern "C" {
int ref_edge_create(int levels, double hyst_low_factor, int normalize_output, void** out) {
    REF_TRY
    *out = new RefEdge((size_t)levels, hyst_low_factor, normalize_output != 0);
    REF_CATCH
}
int ref_edge_destroy(void* h) { delete (RefEdge*)h; return 0; }
int ref_edge_apply_threshold(void* h, const unsigned char* img, int w, int hh, int c, unsigned char* out, double thr) {
    REF_TRY
    cv::Mat m;
    ((RefEdge*)h)->apply_threshold(cv::Mat(hh, w, CV_8UC(c), (void*)img), m, thr);
    if(m.type() != CV_8UC1 || m.rows != hh || m.cols != w) throw std::runtime_error("unexpected edge mask type");
    Bytes t; put_mat(m, t); std::memcpy(out, t.data(), t.size());
    REF_CATCH
}
int ref_edge_apply(void* h, const unsigned char* img, int w, int hh, int c, unsigned char* out) {
    REF_TRY
    cv::Mat m;
    ((RefEdge*)h)->apply(cv::Mat(hh, w, CV_8UC(c), (void*)img), m);
    if(m.type() != CV_8UC1 || m.rows != hh || m.cols != w) throw std::runtime_error("unexpected confidence map type");
    Bytes t; put_mat(m, t); std::memcpy(out, t.data(), t.size());
    REF_CATCH
}
/// the detector's persistent buffers as they are: which = 0 gradient map (4 bytes per cell, padded by 2 on every side), 1 edge mask (padded)
int ref_edge_raw(void* h, int which, unsigned char* out, size_t* bytes) {
    REF_TRY
    const lv::aligned_vector<uchar,32>& v = which == 0 ? ((RefEdge*)h)->grad() : ((RefEdge*)h)->edge();
    if(!out) { *bytes = v.size(); return 0; }
    if(*bytes != v.size()) throw std::runtime_error("size mismatch for the edge detector buffer");
    std::memcpy(out, v.data(), v.size());
    REF_CATCH
}
} // extern "C"

