// ORACLE — TEST INFRASTRUCTURE ONLY (see lvo_common.hpp header).
// C entry points (ctypes) over the CPU restatement; same shape as the product C-ABI in include/litiv_b200.h
// so tests can drive both with one harness.
#include "lvo_common.hpp"
#include "lvo_subsense.hpp"
#ifdef LVO_WITH_PAWCS
#include "lvo_pawcs.hpp"
#endif
#include "lvo_vibe.hpp"
#include "lvo_pbas.hpp"
#include "lvo_edge_lbsp.hpp"
#include <map>
#include <chrono>

using namespace lvo;

namespace {
thread_local std::string g_err;

struct Handle {
    int algo;
    SuBSENSE sub;
    LOBSTER lob;
#ifdef LVO_WITH_PAWCS
    PAWCS paw;
#endif
    BgsBase& base() {
#ifdef LVO_WITH_PAWCS
        if(algo == 2) return paw;
#endif
        return algo == 1 ? (BgsBase&)sub : (BgsBase&)lob;
    }
};

struct Buf { void* ptr; size_t bytes; };

bool find_buf(Handle* h, const std::string& n, Buf& b) {
#define VB(name, vec) if(n == name) { b.ptr = (void*)(vec).data(); b.bytes = (vec).size() * sizeof((vec)[0]); return true; }
    BgsBase& base = h->base();
    VB("roi", base.roi) VB("lastfg", base.last_fg) VB("lastcolor", base.last_color) VB("lastdesc", base.last_desc)
    if(n == "lut") { b.ptr = base.lut; b.bytes = 256; return true; }
    if(h->algo == 1) {
        SuBSENSE& s = h->sub;
        VB("bg_color", s.bg_color) VB("bg_desc", s.bg_desc)
        VB("T", s.T) VB("R", s.R) VB("v", s.V) VB("Dlast", s.Dlast) VB("DminLT", s.DminLT) VB("DminST", s.DminST)
        VB("rawLT", s.rawLT) VB("rawST", s.rawST) VB("finLT", s.finLT) VB("finST", s.finST)
        VB("dsLT", s.dsLT) VB("dsST", s.dsST)
        VB("unstable", s.unstable) VB("blinks", s.blinks) VB("lastraw", s.last_raw) VB("lastrawblink", s.last_raw_blink)
        VB("dilinv", s.dil_inv) VB("rawmask", s.raw_mask)
    } else if(h->algo == 0) {
        LOBSTER& s = h->lob;
        VB("bg_color", s.bg_color) VB("bg_desc", s.bg_desc) VB("rawmask", s.raw_mask)
    }
#ifdef LVO_WITH_PAWCS
    else if(h->algo == 2) { if(h->paw.find_buf(n, b.ptr, b.bytes)) return true; }
#endif
#undef VB
    return false;
}
} // namespace

#define LVO_TRY try {
#define LVO_CATCH } catch(const std::exception& e) { g_err = e.what(); return 1; } return 0;

extern "C" {

const char* lvo_last_error() { return g_err.c_str(); }

int lvo_create(int algo, const Params* p_or_null, int mode, uint64_t seed, void** out) {
    LVO_TRY
    if(algo < 0 || algo > 2) throw std::runtime_error("unknown algorithm id");
#ifndef LVO_WITH_PAWCS
    if(algo == 2) throw std::runtime_error("PAWCS oracle not built");
#endif
    Handle* h = new Handle();
    h->algo = algo;
    BgsBase& b = h->base();
    if(p_or_null) b.P = *p_or_null;
    else {
        if(algo == 0) { b.P.desc_dist_threshold = 4; b.P.color_dist_threshold = 30; b.P.n_samples = 35; b.P.n_required = 2; }
        if(algo == 2) { b.P.desc_dist_threshold = 2; b.P.color_dist_threshold = 20; b.P.n_samples = 50; b.P.n_required = 0; }
    }
    if(b.P.n_samples <= 0 || b.P.n_required > b.P.n_samples) throw std::runtime_error("algo cannot require more sample matches than sample count in model");
    if(b.P.rel_lbsp_threshold < 0) throw std::runtime_error("relative threshold for LBSP features must be non-negative");
    b.mode = (Mode)mode; b.seed = seed;
    b.grand.srand((unsigned)seed);
    *out = h;
    LVO_CATCH
}
int lvo_destroy(void* h) { delete (Handle*)h; return 0; }

int lvo_initialize(void* hv, const uint8_t* img, int w, int h, int c, const uint8_t* roi) {
    LVO_TRY
    Handle* H = (Handle*)hv;
    if(H->algo == 1) H->sub.initialize(img, w, h, c, roi);
    else if(H->algo == 0) H->lob.initialize(img, w, h, c, roi);
#ifdef LVO_WITH_PAWCS
    else H->paw.initialize(img, w, h, c, roi);
#endif
    LVO_CATCH
}
int lvo_apply(void* hv, const uint8_t* img, uint8_t* mask, double lr) {
    LVO_TRY
    Handle* H = (Handle*)hv;
    if(H->algo == 1) H->sub.apply(img, mask, lr);
    else if(H->algo == 0) H->lob.apply(img, mask, lr);
#ifdef LVO_WITH_PAWCS
    else H->paw.apply(img, mask, lr);
#endif
    LVO_CATCH
}
int lvo_refresh_model(void* hv, float frac, int force_fg) {
    LVO_TRY
    Handle* H = (Handle*)hv;
    if(H->algo == 1) H->sub.refresh_model(frac, force_fg != 0);
    else if(H->algo == 0) H->lob.refresh_model(frac, force_fg != 0);
    else throw std::runtime_error("use lvo_pawcs_refresh_model");
    LVO_CATCH
}
int lvo_pawcs_refresh_model(void* hv, uint64_t base_occ, float decr_frac, int force_fg) {
    LVO_TRY
    Handle* H = (Handle*)hv;
#ifdef LVO_WITH_PAWCS
    if(H->algo == 2) { H->paw.refresh_model((size_t)base_occ, decr_frac, force_fg != 0); return 0; }
#endif
    (void)H; (void)base_occ; (void)decr_frac; (void)force_fg;
    throw std::runtime_error("not a PAWCS instance");
    LVO_CATCH
}
int lvo_get_background_image(void* hv, uint8_t* out) {
    LVO_TRY
    Handle* H = (Handle*)hv;
    if(!H->base().initialized) throw std::runtime_error("algo must be initialized first");
    if(H->algo == 1) H->sub.get_background_image(out);
    else if(H->algo == 0) H->lob.get_background_image(out);
#ifdef LVO_WITH_PAWCS
    else H->paw.get_background_image(out);
#endif
    LVO_CATCH
}
int lvo_get_background_descriptors_image(void* hv, uint16_t* out) {
    LVO_TRY
    Handle* H = (Handle*)hv;
    if(!H->base().initialized) throw std::runtime_error("algo must be initialized first");
    if(H->algo == 1) H->sub.get_background_desc_image(out);
    else if(H->algo == 0) H->lob.get_background_desc_image(out);
#ifdef LVO_WITH_PAWCS
    else H->paw.get_background_desc_image(out);
#endif
    LVO_CATCH
}
int lvo_set_auto_model_reset(void* hv, int v) { ((Handle*)hv)->base().auto_reset = v != 0; return 0; }

int lvo_state_size(void* hv, const char* name, size_t* bytes) {
    LVO_TRY
    Buf b;
    if(std::string(name) == "scalars") { *bytes = 16 * sizeof(double); return 0; }
    if(!find_buf((Handle*)hv, name, b)) throw std::runtime_error(std::string("unknown state buffer: ") + name);
    *bytes = b.bytes;
    LVO_CATCH
}
/// "scalars" = 16 doubles: frame_idx, frames_since_reset, reset_cooldown, auto_reset, lr_scaling, use3x3, median_k,
/// t_lower, t_upper, last_nonzero_ratio, roi_count, orig_roi_count, refresh_epoch, 0, 0, 0
int lvo_state_get(void* hv, const char* name, void* out, size_t bytes) {
    LVO_TRY
    Handle* H = (Handle*)hv;
    if(std::string(name) == "scalars") {
        if(bytes != 16 * sizeof(double)) throw std::runtime_error("bad size for scalars");
        double* d = (double*)out; BgsBase& b = H->base();
        std::memset(d, 0, bytes);
        d[0] = (double)b.frame_idx; d[1] = (double)b.frames_since_reset; d[2] = (double)b.reset_cooldown; d[3] = b.auto_reset;
        d[10] = (double)b.roi_count; d[11] = (double)b.orig_roi_count; d[12] = b.refresh_epoch;
        if(H->algo == 1) { SuBSENSE& s = H->sub; d[4] = s.lr_scaling; d[5] = s.use3x3; d[6] = s.median_k; d[7] = s.t_lower; d[8] = s.t_upper; d[9] = s.last_nonzero_ratio; }
#ifdef LVO_WITH_PAWCS
        if(H->algo == 2) H->paw.get_scalars(d);
#endif
        return 0;
    }
    Buf b;
    if(!find_buf(H, name, b)) throw std::runtime_error(std::string("unknown state buffer: ") + name);
    if(b.bytes != bytes) throw std::runtime_error(std::string("size mismatch for state buffer ") + name);
    std::memcpy(out, b.ptr, bytes);
    LVO_CATCH
}
int lvo_state_set(void* hv, const char* name, const void* in, size_t bytes) {
    LVO_TRY
    Handle* H = (Handle*)hv;
    if(std::string(name) == "scalars") {
        if(bytes != 16 * sizeof(double)) throw std::runtime_error("bad size for scalars");
        const double* d = (const double*)in; BgsBase& b = H->base();
        b.frame_idx = (size_t)d[0]; b.frames_since_reset = (size_t)d[1]; b.reset_cooldown = (size_t)d[2]; b.auto_reset = d[3] != 0;
        b.refresh_epoch = (uint32_t)d[12];
        if(H->algo == 1) { SuBSENSE& s = H->sub; s.lr_scaling = d[4] != 0; s.use3x3 = d[5] != 0; s.median_k = (int)d[6]; s.t_lower = (float)d[7]; s.t_upper = (float)d[8]; s.last_nonzero_ratio = (float)d[9]; }
#ifdef LVO_WITH_PAWCS
        if(H->algo == 2) H->paw.set_scalars(d);
#endif
        return 0;
    }
    Buf b;
    if(!find_buf(H, name, b)) throw std::runtime_error(std::string("unknown state buffer: ") + name);
    if(b.bytes != bytes) throw std::runtime_error(std::string("size mismatch for state buffer ") + name);
    std::memcpy(b.ptr, in, bytes);
    LVO_CATCH
}
int lvo_get_scan_hist(void* hv, uint64_t out[64]) {
    const Stats& s = ((Handle*)hv)->base().stats;
    for(int i = 0; i < 64; ++i) out[i] = s.scan_hist[i];
    return 0;
}
/// stats: roi_px, samples_scanned, sample_writes, fg_px, frames (accumulated since initialize)
int lvo_get_stats(void* hv, uint64_t out[5]) {
    const Stats& s = ((Handle*)hv)->base().stats;
    out[0] = s.roi_px; out[1] = s.samples_scanned; out[2] = s.sample_writes; out[3] = s.fg_px; out[4] = s.frames;
    return 0;
}

int lvo_lbsp_compute(const uint8_t* img, const uint8_t* ref_or_null, int w, int h, int c, int use_rel, float rel, int thr, uint16_t* out) {
    LVO_TRY
    if(!img || (c != 1 && c != 3)) throw std::runtime_error("input image must be non-empty, continuous, and of type 8UC1/8UC3");
    if(w < 5 || h < 5) throw std::runtime_error("input image size is too small to compute descriptors with current patch size");
    if(use_rel && rel < 0) throw std::runtime_error("lbsp internal relative threshold must be non-negative");
    lbsp_compute_dense(img, ref_or_null, w, h, c, use_rel != 0, rel, (size_t)thr, out);
    LVO_CATCH
}

// lv::BinClassif::accumulate(oClassif, oGT, oROI) -- modules/datasets/src/metrics.cpp:21-61; label values metrics.hpp:23-27.
// counters = TP, TN, FP, FN, SE, DC (BinClassif::CountersList, metrics.hpp:40-48), added to.
int lvo_binclassif(const uint8_t* classif, const uint8_t* gt_or_null, const uint8_t* roi_or_null, int w, int h, uint64_t counters[6]) {
    LVO_TRY
    if(!classif || w < 1 || h < 1) throw std::runtime_error("binary classifier results must be non-empty and of type 8UC1");
    const size_t n = (size_t)w * h;
    if(!gt_or_null) { counters[5] += n; return 0; }                       // :26-29
    for(size_t i = 0; i < n; ++i) {
        const uint8_t g = gt_or_null[i], in = classif[i];
        if(g != 85 && g != 170 && (!roi_or_null || roi_or_null[i] != 0)) { // :37-39 (out of scope, unknown, ROI negative)
            if(in == 255) { if(g == 255) ++counters[0]; else ++counters[2]; }   // TP / FP (:40-45)
            else          { if(g == 255) ++counters[3]; else ++counters[1]; }   // FN / TN (:46-51)
            if(g == 50 && in == 255) ++counters[4];                              // shadow error (:52-55)
        } else ++counters[5];                                              // :57-58
    }
    LVO_CATCH
}

// --- helper entry points used by the "not gpu" tests to pin the helpers against the reference's known answers
int lvo_glibc_rand_seq(unsigned seed, int n, int* out) { GlibcRand g(seed); for(int i = 0; i < n; ++i) out[i] = g.next(); return 0; }
int lvo_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) { philox4x32_10(ctr, key, out); return 0; }
int lvo_philox_draw(uint64_t seed, uint32_t frame, uint32_t pixel, uint32_t site, uint32_t domain) { return philox_draw(seed, frame, pixel, site, domain); }
int lvo_l1dist3_u8(const uint8_t* a, const uint8_t* b) { return L1dist_arr_u8<3>(a, b); }
uint64_t lvo_cdist3(const uint8_t* a, const uint8_t* b) { return cdist_u8<3>(a, b); }
uint64_t lvo_cdist4(const uint8_t* a, const uint8_t* b) { return cdist_u8<4>(a, b); }
uint64_t lvo_cdist2(const uint8_t* a, const uint8_t* b) { return cdist_u8<2>(a, b); }
int lvo_sample_pos_7x7(int rnd, int ox, int oy, int border, int W, int H, int* xy) { sample_pos_7x7(rnd, xy[0], xy[1], ox, oy, border, W, H); return 0; }
int lvo_neighbor_pos(int five, int rnd, int ox, int oy, int border, int W, int H, int* xy) {
    if(five) neighbor_pos_5x5(rnd, xy[0], xy[1], ox, oy, border, W, H); else neighbor_pos_3x3(rnd, xy[0], xy[1], ox, oy, border, W, H);
    return 0;
}
int lvo_morph_rect(const uint8_t* src, uint8_t* dst, int W, int H, int r, int dilate) { morph_rect(src, dst, W, H, r, dilate != 0); return 0; }
int lvo_median_binary(const uint8_t* src, uint8_t* dst, int W, int H, int k) { median_binary(src, dst, W, H, k); return 0; }
int lvo_floodfill_origin(uint8_t* img, int W, int H) { floodfill_from_origin(img, W, H); return 0; }
int lvo_resize_area_general(const uint8_t* src, int W, int H, int C, int dw, int dh, uint8_t* dst) { resize_area_general(src, W, H, C, dw, dh, dst); return 0; }
int lvo_resize_area_exact(const uint8_t* src, int W, int H, int C, int s, uint8_t* dst) { resize_area_exact(src, W, H, C, s, dst); return 0; }
int lvo_lbsp_threshold(const uint8_t* vals, int ref, int t, int scalar) { return scalar ? lbsp_threshold_scalar(vals, (uchar)ref, (uchar)t) : lbsp_threshold(vals, (uchar)ref, (uchar)t); }
int lvo_build_lut(int C, float rel, int off, uint8_t* lut) { build_lbsp_lut(C, rel, (size_t)off, lut); return 0; }

/// timing helper for bench.py's cpu_baseline: run `n` frames back-to-back from a [n][H][W][C] buffer,
/// return seconds spent inside apply() only.
double lvo_apply_sequence(void* hv, const uint8_t* frames, int n, size_t frame_bytes, uint8_t* last_mask, const double* lrs) {
    double total = 0;
    for(int i = 0; i < n; ++i) {
        auto t0 = std::chrono::steady_clock::now();
        if(lvo_apply(hv, frames + (size_t)i * frame_bytes, last_mask, lrs[i]) != 0) return -1.0;
        total += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    }
    return total;
}

int lvo_lbsp_gradient(const uint8_t* img, int w, int h, int c, uint8_t* out) {
    LVO_TRY
    if(!img || c < 1 || c > 4 || w < 5 || h < 5) throw std::runtime_error("input image must be non-empty, 8UC1 .. 8UC4 and at least 5x5");
    lbsp_gradient_dense(img, w, h, c, out);
    LVO_CATCH
}

// --- EdgeDetectorLBSP (imgproc/src/EdgeDetectorLBSP.cpp): the checker of lvb_edge_* 
int lvo_edge_create(int levels, double hyst_low_factor, void** out) {
    LVO_TRY
    if(levels < 1) throw std::runtime_error("number of pyramid levels must be positive");
    if(!(hyst_low_factor > 0 && hyst_low_factor < 1)) throw std::runtime_error("lower hysteresis threshold factor must be between 0 and 1");
    EdgeDetectorLBSP* e = new EdgeDetectorLBSP();
    e->n_levels = levels; e->hyst_low_factor = hyst_low_factor;
    *out = e;
    LVO_CATCH
}
int lvo_edge_destroy(void* h) { delete (EdgeDetectorLBSP*)h; return 0; }
int lvo_edge_set_normalize(void* h, int on) { ((EdgeDetectorLBSP*)h)->normalize_output = on != 0; return 0; }
int lvo_normalize_minmax_u8(uint8_t* buf, size_t n) { EdgeDetectorLBSP::normalize_minmax_u8(buf, n); return 0; }
int lvo_edge_apply_threshold(void* h, const uint8_t* img, int w, int hh, int c, uint8_t* out, double thr) { LVO_TRY ((EdgeDetectorLBSP*)h)->apply_threshold(img, w, hh, c, out, thr); LVO_CATCH }
int lvo_edge_apply(void* h, const uint8_t* img, int w, int hh, int c, uint8_t* out) { LVO_TRY ((EdgeDetectorLBSP*)h)->apply(img, w, hh, c, out); LVO_CATCH }
/// gradient map of the latest pass without its padding: [H][W][4] = gradX, gradY, magnitude (min over the scales), pad
/// the detector's persistent buffers as they are: which = 0 gradient map (4 bytes per cell, padded by 2 on every side), 1 edge mask (padded)
int lvo_edge_raw(void* h, int which, uint8_t* out, size_t* bytes) {
    LVO_TRY
    const std::vector<uchar>& v = which == 0 ? ((EdgeDetectorLBSP*)h)->grad : ((EdgeDetectorLBSP*)h)->edge;
    if(!out) { *bytes = v.size(); return 0; }
    if(*bytes != v.size()) throw std::runtime_error("size mismatch for the edge detector buffer");
    std::memcpy(out, v.data(), v.size());
    LVO_CATCH
}
int lvo_edge_gradient_map(void* h, int w, int hh, uint8_t* out) {
    LVO_TRY
    EdgeDetectorLBSP* e = (EdgeDetectorLBSP*)h;
    if(e->grad.size() != (size_t)(w + 4) * (hh + 4) * 4) throw std::runtime_error("no pass of that size has run");
    for(int r = 0; r < hh; ++r) std::memcpy(out + (size_t)r * w * 4, e->grad.data() + ((size_t)(r + 2) * (w + 4) + 2) * 4, (size_t)w * 4);
    LVO_CATCH
}

// --- ViBe (video/src/BackgroundSubtractorViBe.cpp): separate handle type, the class is a plain cv::BackgroundSubtractor (no ROI)
int lvo_vibe_create(int model_channels, int color_dist_threshold, int n_samples, int n_required, int mode, uint64_t seed, void** out) {
    LVO_TRY
    if(model_channels != 1 && model_channels != 3) throw std::runtime_error("ViBe model must have 1 or 3 channels");
    if(n_samples <= 0 || n_required > n_samples) throw std::runtime_error("algo cannot require more sample matches than sample count in model");
    ViBe* v = new ViBe();
    v->model_channels = model_channels; v->color_dist_threshold = color_dist_threshold; v->n_samples = n_samples; v->n_required = n_required;
    v->mode = (Mode)mode; v->seed = seed; v->grand.srand((unsigned)seed);
    *out = v;
    LVO_CATCH
}
int lvo_vibe_destroy(void* h) { delete (ViBe*)h; return 0; }
int lvo_vibe_initialize(void* h, const uint8_t* img, int w, int hh, int c) { LVO_TRY ((ViBe*)h)->initialize(img, w, hh, c); LVO_CATCH }
int lvo_vibe_apply(void* h, const uint8_t* img, int c, uint8_t* mask, double lr) { LVO_TRY ((ViBe*)h)->apply(img, c, mask, lr); LVO_CATCH }
int lvo_vibe_get_background_image(void* h, uint8_t* out) { LVO_TRY ((ViBe*)h)->get_background_image(out); LVO_CATCH }
/// model samples in the reference's layout [N][H][W][C]
int lvo_vibe_model(void* h, uint8_t* inout, size_t bytes, int set) {
    LVO_TRY
    ViBe* v = (ViBe*)h;
    if(bytes != v->bg.size()) throw std::runtime_error("size mismatch for the ViBe model");
    if(set) std::memcpy(v->bg.data(), inout, bytes); else std::memcpy(inout, v->bg.data(), bytes);
    LVO_CATCH
}
int lvo_vibe_set_frame(void* h, uint64_t frame_idx) { ((ViBe*)h)->frame_idx = (size_t)frame_idx; return 0; }
int lvo_vibe_get_stats(void* h, uint64_t out[5]) {
    const Stats& s = ((ViBe*)h)->stats;
    out[0] = s.roi_px; out[1] = s.samples_scanned; out[2] = s.sample_writes; out[3] = s.fg_px; out[4] = s.frames;
    return 0;
}
int lvo_vibe_match(int model_channels, int thr, const uint8_t* a, const uint8_t* b) {
    ViBe v; v.model_channels = model_channels; v.color_dist_threshold = thr;
    return v.matches(a, b) ? 1 : 0;
}

// --- PBAS (video/src/BackgroundSubtractorPBAS.cpp): separate handle type, like ViBe
int lvo_pbas_create(int model_channels, int color_dist_threshold, float update_rate, int n_samples, int n_required, int mode, uint64_t seed, void** out) {
    LVO_TRY
    if(model_channels != 1 && model_channels != 3) throw std::runtime_error("PBAS model must have 1 or 3 channels");
    if(n_samples <= 0 || n_required > n_samples) throw std::runtime_error("algo cannot require more sample matches than sample count in model");
    if(!(update_rate > 0 && update_rate <= 255)) throw std::runtime_error("default update rate must be in ]0,255]");
    PBAS* v = new PBAS();
    v->model_channels = model_channels; v->color_dist_threshold = color_dist_threshold; v->default_update_rate = update_rate;
    v->n_samples = n_samples; v->n_required = n_required;
    v->mode = (Mode)mode; v->seed = seed; v->grand.srand((unsigned)seed);
    *out = v;
    LVO_CATCH
}
int lvo_pbas_destroy(void* h) { delete (PBAS*)h; return 0; }
int lvo_pbas_initialize(void* h, const uint8_t* img, int w, int hh, int c) { LVO_TRY ((PBAS*)h)->initialize(img, w, hh, c); LVO_CATCH }
int lvo_pbas_apply(void* h, const uint8_t* img, int c, uint8_t* mask, double lr) { LVO_TRY ((PBAS*)h)->apply(img, c, mask, lr); LVO_CATCH }
int lvo_pbas_get_background_image(void* h, uint8_t* out) { LVO_TRY ((PBAS*)h)->get_background_image(out); LVO_CATCH }
/// named state: bg_color / bg_grad [N][H][W][C] u8, R / T / meanmin [H][W] f32, rawmask / lastgrad u8, scalars = {frame_idx, former_mean_grad_dist} f64
int lvo_pbas_state(void* h, const char* name, void* inout, size_t bytes, int set) {
    LVO_TRY
    PBAS* v = (PBAS*)h;
    const std::string n(name);
    void* ptr = nullptr; size_t sz = 0;
    double sc[2] = {(double)v->frame_idx, (double)v->former_mean_grad_dist};
    if(n == "bg_color") { ptr = v->bg_color.data(); sz = v->bg_color.size(); }
    else if(n == "bg_grad") { ptr = v->bg_grad.data(); sz = v->bg_grad.size(); }
    else if(n == "R") { ptr = v->R.data(); sz = v->R.size() * 4; }
    else if(n == "T") { ptr = v->T.data(); sz = v->T.size() * 4; }
    else if(n == "meanmin") { ptr = v->meanmin.data(); sz = v->meanmin.size() * 4; }
    else if(n == "rawmask") { ptr = v->raw_mask.data(); sz = v->raw_mask.size(); }
    else if(n == "lastgrad") { ptr = v->last_grad.data(); sz = v->last_grad.size(); }
    else if(n == "scalars") { ptr = sc; sz = sizeof(sc); }
    else throw std::runtime_error("unknown state buffer: " + n);
    if(bytes != sz) throw std::runtime_error("size mismatch for state buffer " + n);
    if(set) { std::memcpy(ptr, inout, sz); if(n == "scalars") { v->frame_idx = (size_t)sc[0]; v->former_mean_grad_dist = (float)sc[1]; } }
    else std::memcpy(inout, ptr, sz);
    LVO_CATCH
}
int lvo_pbas_get_stats(void* h, uint64_t out[5]) {
    const Stats& s = ((PBAS*)h)->stats;
    out[0] = s.roi_px; out[1] = s.samples_scanned; out[2] = s.sample_writes; out[3] = s.fg_px; out[4] = s.frames;
    return 0;
}
int lvo_pbas_gradient_image(const uint8_t* img, int w, int h, int c, uint8_t* out) { LVO_TRY pbas_gradient_image(img, w, h, c, out); LVO_CATCH }

} // extern "C"
