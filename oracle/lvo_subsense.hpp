// ORACLE — TEST INFRASTRUCTURE ONLY (see lvo_common.hpp header).
// CPU restatement of BackgroundSubtractorSuBSENSE (video/src/BackgroundSubtractorSuBSENSE.cpp) and
// BackgroundSubtractorLOBSTER (video/src/BackgroundSubtractorLOBSTER.cpp:410-620), plus their common
// bases (video/src/BackgroundSubtractionUtils.cpp:70-155, video/src/BackgroundSubtractorLBSP.cpp:21-65).
//
// Two run modes:
//   MODE_REFERENCE : the reference's semantics — raster order, one sequential libc-rand() stream whose
//                    draw count depends on control flow, neighbour writes visible to later pixels of
//                    the same frame, ghost rule reading live neighbour maps (quirks Q5/Q6).
//   MODE_SNAPSHOT  : the deterministic parallel semantics the GPU implements — every pixel reads the
//                    frame-start state; own-slot writes apply at once (only the owner reads its own
//                    samples), neighbour writes are queued and applied after the pixel pass in raster
//                    order of their SOURCE pixel (last writer wins); draws are Philox4x32-10 indexed by
//                    (seed; frame, pixel, site) so they do not depend on control flow.
#pragma once
#include "lvo_common.hpp"

namespace lvo {

struct Params {
    float rel_lbsp_threshold = 0.333f;
    int lbsp_threshold_offset = 0;
    int desc_dist_threshold = 3;   // SuBSENSE/PAWCS: offset; LOBSTER: absolute threshold
    int color_dist_threshold = 30; // SuBSENSE/PAWCS: minimum (R0); LOBSTER: absolute threshold
    int n_samples = 50;
    int n_required = 2;
    int n_samples_for_moving_avgs = 100;
    int n_global_words = 25; // PAWCS only
    int median_blur_kernel_size = 9;
};

struct Stats { // instrumentation for the roofline's algorithmic-bytes figure
    uint64_t roi_px = 0, samples_scanned = 0, sample_writes = 0, fg_px = 0, frames = 0;
    uint64_t scan_hist[64] = {}; // histogram of the per-pixel scan depth (last bin: >= 63)
};

struct BgsBase {
    Params P;
    Mode mode = MODE_REFERENCE;
    uint64_t seed = 0;
    GlibcRand grand;
    int W = 0, H = 0, C = 0;
    size_t npx = 0;
    std::vector<uchar> roi;
    size_t orig_roi_count = 0, roi_count = 0;
    size_t frame_idx = 0, frames_since_reset = 0, reset_cooldown = 0;
    bool initialized = false, auto_reset = true;
    std::vector<uchar> last_fg, last_color;
    std::vector<ushort> last_desc;
    uchar lut[256];
    uint32_t refresh_epoch = 0;
    Stats stats;
    virtual ~BgsBase() {}

    int draw_ref() { return grand.next(); }

    /// IIBackgroundSubtractor::initialize_common (BackgroundSubtractionUtils.cpp:70-155) +
    /// IBackgroundSubtractorLBSP_::initialize_common (BackgroundSubtractorLBSP.cpp:21-65, quirk Q4)
    void initialize_common(const uchar* img, int w, int h, int c, const uchar* roi_or_null) {
        if(!img || w <= 0 || h <= 0) throw std::runtime_error("provided image for initialization must be non-empty, continuous, and of type 8UC1/3/4");
        if(c != 1 && c != 3) throw std::runtime_error("only 8UC1 and 8UC3 images are supported");
        if(w < 5 || h < 5) throw std::runtime_error("image too small for the LBSP pattern");
        W = w; H = h; C = c; npx = (size_t)W * H;
        build_roi(roi_or_null, W, H, 2, roi, orig_roi_count, roi_count);
        initialized = false;
        frame_idx = 0; frames_since_reset = 0; reset_cooldown = 0; refresh_epoch = 0;
        last_fg.assign(npx, 0);
        last_color.assign(npx * C, 0);
        for(size_t p = 0; p < npx; ++p) if(roi[p]) for(int k = 0; k < C; ++k) last_color[p * C + k] = img[p * C + k];
        last_desc.assign(npx * C, 0);
        build_lbsp_lut(C, P.rel_lbsp_threshold, (size_t)P.lbsp_threshold_offset, lut);
        for(int y = 0; y < H; ++y) for(int x = 0; x < W; ++x) {
            const size_t p = (size_t)y * W + x;
            if(roi[p] && x > 2 && y > 2 && x < W - 2 && y < H - 2) { // strict '>' : Q4
                for(int k = 0; k < C; ++k) {
                    uchar vals[16];
                    lbsp_lookup(img, W, C, x, y, k, vals);
                    last_desc[p * C + k] = lbsp_threshold(vals, img[p * C + k], lut[img[p * C + k]]);
                }
            }
        }
        stats = Stats();
    }
};

// =================================================================================================
// SuBSENSE
// =================================================================================================
struct SuBSENSE : BgsBase {
    // SuBSENSE.cpp:26-46
    static constexpr float GHOSTDET_D_MAX = 0.010f, GHOSTDET_S_MIN = 0.995f;
    static constexpr float FEEDBACK_R_VAR = 0.01f, FEEDBACK_V_INCR = 1.0f, FEEDBACK_V_DECR = 0.1f;
    static constexpr float FEEDBACK_T_DECR = 0.25f, FEEDBACK_T_INCR = 0.5f, FEEDBACK_T_LOWER = 2.0f, FEEDBACK_T_UPPER = 256.0f;
    static constexpr float UNSTABLE_REG_RATIO_MIN = 0.1f, UNSTABLE_REG_RDIST_MIN = 3.0f;
    static constexpr float LBSPDESC_NONZERO_RATIO_MIN = 0.1f, LBSPDESC_NONZERO_RATIO_MAX = 0.5f;

    bool lr_scaling = true, use3x3 = true;
    int median_k = 9;
    float t_lower = 2.0f, t_upper = 256.0f, last_nonzero_ratio = 0.0f;
    int dsW = 0, dsH = 0;
    std::vector<float> T, R, V, Dlast, DminLT, DminST, rawLT, rawST, finLT, finST, dsLT, dsST;
    std::vector<uchar> unstable, blinks, last_raw, last_raw_blink, dil_inv, raw_mask, ds_frame;
    std::vector<uchar> bg_color;  // [N][H*W*C]
    std::vector<ushort> bg_desc;  // [N][H*W*C]

    uchar* bgc(int s) { return bg_color.data() + (size_t)s * npx * C; }
    ushort* bgd(int s) { return bg_desc.data() + (size_t)s * npx * C; }

    /// SuBSENSE.cpp:80-105
    void refresh_model(float frac, bool force_fg) {
        if(!initialized) throw std::runtime_error("algo must be initialized first");
        if(!(frac > 0.0f && frac <= 1.0f)) throw std::runtime_error("model refresh must be given as a non-null fraction");
        const size_t N = (size_t)P.n_samples;
        const size_t n_refresh = frac < 1.0f ? (size_t)(frac * N) : N;
        const uint32_t epoch = refresh_epoch++;
        size_t start = 0;
        if(frac < 1.0f) start = (size_t)(mode == MODE_REFERENCE ? draw_ref() : philox_draw(seed, epoch, 0, 0, DOM_REFRESH_START)) % N;
        for(size_t p = 0; p < npx; ++p) {
            if(!roi[p]) continue;
            if(force_fg || !last_fg[p]) {
                const int ox = (int)(p % W), oy = (int)(p / W);
                for(size_t s = start; s < start + n_refresh; ++s) {
                    const size_t rs = s % N;
                    const int rnd = mode == MODE_REFERENCE ? draw_ref() : philox_draw(seed, epoch, (uint32_t)p, (uint32_t)rs, DOM_REFRESH);
                    int sx, sy;
                    sample_pos_7x7(rnd, sx, sy, ox, oy, 2, W, H);
                    const size_t sp = (size_t)sy * W + sx;
                    if(force_fg || !last_fg[sp]) {
                        for(int c = 0; c < C; ++c) {
                            bgc((int)rs)[p * C + c] = last_color[sp * C + c];
                            bgd((int)rs)[p * C + c] = last_desc[sp * C + c];
                        }
                    }
                }
            }
        }
    }

    /// SuBSENSE.cpp:107-186
    void initialize(const uchar* img, int w, int h, int c, const uchar* roi_or_null) {
        initialize_common(img, w, h, c, roi_or_null);
        last_nonzero_ratio = 0.0f;
        const int tot = W * H;
        const int qvga = 320 * 240;
        if(orig_roi_count >= npx / 2 && tot >= qvga) {
            lr_scaling = true; auto_reset = true;
            use3x3 = !(tot > qvga * 2);
            const int rawk = std::min((int)std::floor((float)tot / qvga + 0.5f) + P.median_blur_kernel_size, 14);
            median_k = (rawk % 2) ? rawk : rawk - 1;
            t_lower = FEEDBACK_T_LOWER; t_upper = FEEDBACK_T_UPPER;
        } else {
            lr_scaling = false; auto_reset = false; use3x3 = true;
            median_k = P.median_blur_kernel_size;
            t_lower = FEEDBACK_T_LOWER * 2; t_upper = FEEDBACK_T_UPPER * 2;
        }
        T.assign(npx, t_lower); R.assign(npx, 1.0f); V.assign(npx, 10.0f);
        Dlast.assign(npx, 0.f); DminLT.assign(npx, 0.f); DminST.assign(npx, 0.f);
        rawLT.assign(npx, 0.f); rawST.assign(npx, 0.f); finLT.assign(npx, 0.f); finST.assign(npx, 0.f);
        dsW = W / 8; dsH = H / 8;
        dsLT.assign((size_t)dsW * dsH * C, 0.f); dsST.assign((size_t)dsW * dsH * C, 0.f);
        ds_frame.assign((size_t)dsW * dsH * C, 0);
        unstable.assign(npx, 0); blinks.assign(npx, 0); last_raw.assign(npx, 0); last_raw_blink.assign(npx, 0);
        dil_inv.assign(npx, 0); raw_mask.assign(npx, 0);
        bg_color.assign((size_t)P.n_samples * npx * C, 0);
        bg_desc.assign((size_t)P.n_samples * npx * C, 0);
        initialized = true;
        refresh_model(1.0f, false);
    }

    struct NbWrite { size_t target; int slot; uchar col[3]; ushort desc[3]; };

    /// SuBSENSE.cpp:188-612
    template<int CH> void apply_impl(const uchar* img, uchar* fgmask, double lr_override) {
        const size_t N = (size_t)P.n_samples, REQ = (size_t)P.n_required;
        const size_t minColor = (size_t)P.color_dist_threshold, descOff = (size_t)P.desc_dist_threshold;
        const size_t stabColorOff = minColor / 5, unstabDescOff = descOff;
        const size_t colorRange = CH == 1 ? 255 : 765, descRange = CH == 1 ? 16 : 48;
        std::fill(raw_mask.begin(), raw_mask.end(), 0);
        size_t nonzero_desc = 0;
        ++frame_idx;
        const float aLT = 1.0f / std::min(frame_idx, (size_t)P.n_samples_for_moving_avgs);
        const float aST = 1.0f / std::min(frame_idx, (size_t)P.n_samples_for_moving_avgs / 4);
        const uint32_t fr = (uint32_t)frame_idx;
        std::vector<float> snapDlast, snapRawST;
        std::vector<NbWrite> queue;
        if(mode == MODE_SNAPSHOT) { snapDlast = Dlast; snapRawST = rawST; }
        for(size_t p = 0; p < npx; ++p) {
            if(!roi[p]) continue;
            const int x = (int)(p % W), y = (int)(p / W);
            const uchar* cur = img + p * CH;
            auto draw = [&](uint32_t site) -> size_t {
                return (size_t)(mode == MODE_REFERENCE ? draw_ref() : philox_draw(seed, fr, (uint32_t)p, site, DOM_APPLY));
            };
            size_t minDesc = descRange, minSum = colorRange;
            // thresholds from R(x) and the PREVIOUS unstable flag (:222-223 / :355-359 ; Q3, Q5)
            size_t thrC = (size_t)((R[p] * minColor) - ((!unstable[p]) * stabColorOff));
            if(CH == 1) thrC /= 2;
            const size_t thrD = ((size_t)1 << ((size_t)std::floor(R[p] + 0.5f))) + descOff + (unstable[p] * unstabDescOff);
            const size_t totC = thrC * 3, totD = thrD * 3, scC = totC / 2;
            uchar vals[CH][16];
            ushort intra[CH];
            for(int c = 0; c < CH; ++c) {
                lbsp_lookup(img, W, CH, x, y, c, vals[c]);
                intra[c] = lbsp_threshold(vals[c], cur[c], lut[cur[c]]);
            }
            unstable[p] = (R[p] > UNSTABLE_REG_RDIST_MIN || (rawLT[p] - finLT[p]) > UNSTABLE_REG_RATIO_MIN || (rawST[p] - finST[p]) > UNSTABLE_REG_RATIO_MIN) ? 1 : 0;
            size_t good = 0, s = 0;
            while(good < REQ && s < N) {
                const uchar* bc = bgc((int)s) + p * CH;
                const ushort* bd = bgd((int)s) + p * CH;
                bool ok = true;
                size_t totDesc = 0, totSum = 0;
                if(CH == 1) { // :230-249
                    const size_t cd = L1dist_u8(cur[0], bc[0]);
                    if(cd > thrC) ok = false;
                    else {
                        const size_t dd = ((size_t)hdist16(intra[0], bd[0]) + (size_t)hdist16(lbsp_threshold(vals[0], bc[0], lut[bc[0]]), bd[0])) / 2;
                        if(dd > thrD) ok = false;
                        else {
                            const size_t sum = std::min((dd / 4) * (255 / 16) + cd, (size_t)255);
                            if(sum > thrC) ok = false;
                            else { totDesc = dd; totSum = sum; }
                        }
                    }
                } else { // :368-387
                    for(int c = 0; c < CH && ok; ++c) {
                        const size_t cd = L1dist_u8(cur[c], bc[c]);
                        if(cd > scC) { ok = false; break; }
                        const size_t dd = ((size_t)hdist16(intra[c], bd[c]) + (size_t)hdist16(lbsp_threshold(vals[c], bc[c], lut[bc[c]]), bd[c])) / 2;
                        const size_t sum = std::min((dd / 2) * (255 / 16) + cd, (size_t)255);
                        if(sum > scC) { ok = false; break; }
                        totDesc += dd; totSum += sum;
                    }
                    if(ok && (totDesc > totD || totSum > totC)) ok = false;
                }
                if(ok) {
                    if(minDesc > totDesc) minDesc = totDesc;
                    if(minSum > totSum) minSum = totSum;
                    ++good;
                }
                ++s;
            }
            stats.samples_scanned += s; stats.scan_hist[s < 63 ? s : 63] += 1;
            // :254-255 / :396-397 (Q1: the 3ch L1 wraps in uint8)
            const uchar* lc = last_color.data() + p * CH;
            ushort* ld = last_desc.data() + p * CH;
            size_t lastL1, lastHd = 0;
            if(CH == 1) lastL1 = L1dist_u8(lc[0], cur[0]); else lastL1 = L1dist_arr_u8<CH>(lc, cur);
            for(int c = 0; c < CH; ++c) lastHd += (size_t)hdist16(ld[c], intra[c]);
            const float normLast = ((float)lastL1 / colorRange + (float)lastHd / descRange) / 2;
            Dlast[p] = Dlast[p] * (1.0f - aST) + normLast * aST;
            auto write_sample = [&](size_t target, size_t slot) {
                for(int c = 0; c < CH; ++c) { bgd((int)slot)[target * CH + c] = intra[c]; bgc((int)slot)[target * CH + c] = cur[c]; }
                ++stats.sample_writes;
            };
            if(good < REQ) { // foreground (:256-269 / :398-413)
                const float normMin = std::min(1.0f, ((float)minSum / colorRange + (float)minDesc / descRange) / 2 + (float)(REQ - good) / REQ);
                DminLT[p] = DminLT[p] * (1.0f - aLT) + normMin * aLT;
                DminST[p] = DminST[p] * (1.0f - aST) + normMin * aST;
                rawLT[p] = rawLT[p] * (1.0f - aLT) + aLT;
                rawST[p] = rawST[p] * (1.0f - aST) + aST;
                raw_mask[p] = 255;
                ++stats.fg_px;
                if(reset_cooldown && (draw(0) % (size_t)FEEDBACK_T_LOWER) == 0) {
                    const size_t slot = draw(1) % N;
                    write_sample(p, slot);
                }
            } else { // background (:270-301 / :414-450)
                const float normMin = ((float)minSum / colorRange + (float)minDesc / descRange) / 2;
                DminLT[p] = DminLT[p] * (1.0f - aLT) + normMin * aLT;
                DminST[p] = DminST[p] * (1.0f - aST) + normMin * aST;
                rawLT[p] = rawLT[p] * (1.0f - aLT);
                rawST[p] = rawST[p] * (1.0f - aST);
                const size_t LR = std::isinf(lr_override) ? SIZE_MAX : (lr_override > 0 ? (size_t)std::ceil(lr_override) : (size_t)std::ceil(T[p]));
                if((draw(0) % LR) == 0) {
                    const size_t slot = draw(1) % N;
                    write_sample(p, slot);
                }
                const bool cur3x3 = use3x3 && !unstable[p];
                int nx, ny;
                const int rnb = (int)draw(2);
                if(cur3x3) neighbor_pos_3x3(rnb, nx, ny, x, y, 2, W, H); else neighbor_pos_5x5(rnb, nx, ny, x, y, 2, W, H);
                const size_t n_rand = draw(3);
                const size_t q = (size_t)ny * W + nx;
                const float nbDlast = mode == MODE_REFERENCE ? Dlast[q] : snapDlast[q];
                const float nbRawST = mode == MODE_REFERENCE ? rawST[q] : snapRawST[q];
                if((n_rand % (cur3x3 ? LR : (LR / 2 + 1))) == 0
                   || (nbRawST > GHOSTDET_S_MIN && nbDlast < GHOSTDET_D_MAX && (n_rand % ((size_t)t_lower)) == 0)) {
                    // snapshot mode: the neighbour slot reuses draw site 1 ((d1 / N) % N) so one Philox block serves the whole pixel
                    const size_t slot = mode == MODE_REFERENCE ? draw(4) % N : (draw(1) / N) % N;
                    if(mode == MODE_REFERENCE) write_sample(q, slot);
                    else {
                        NbWrite w; w.target = q; w.slot = (int)slot;
                        for(int c = 0; c < CH; ++c) { w.col[c] = cur[c]; w.desc[c] = intra[c]; }
                        queue.push_back(w);
                    }
                }
            }
            // T(x) (:302-311 / :451-460 ; Q7: the division may give +inf on frame 1, the clamp absorbs it)
            if(last_fg[p] || (std::min(DminLT[p], DminST[p]) < UNSTABLE_REG_RATIO_MIN && raw_mask[p])) {
                if(T[p] < t_upper) T[p] += FEEDBACK_T_INCR / (std::max(DminLT[p], DminST[p]) * V[p]);
            } else if(T[p] > t_lower)
                T[p] -= FEEDBACK_T_DECR * V[p] / std::max(DminLT[p], DminST[p]);
            if(T[p] < t_lower) T[p] = t_lower; else if(T[p] > t_upper) T[p] = t_upper;
            // v(x) (:312-318 / :461-467)
            if(std::max(DminLT[p], DminST[p]) > UNSTABLE_REG_RATIO_MIN && blinks[p]) V[p] += FEEDBACK_V_INCR;
            else if(V[p] > FEEDBACK_V_DECR) {
                V[p] -= last_fg[p] ? FEEDBACK_V_DECR / 4 : unstable[p] ? FEEDBACK_V_DECR / 2 : FEEDBACK_V_DECR;
                if(V[p] < FEEDBACK_V_DECR) V[p] = FEEDBACK_V_DECR;
            }
            // R(x) (:319-325 / :468-474 ; Q7: std::pow(float,int) evaluates in double)
            if((double)R[p] < std::pow((double)(1.0f + std::min(DminLT[p], DminST[p]) * 2), 2.0))
                R[p] += FEEDBACK_R_VAR * (V[p] - FEEDBACK_V_DECR);
            else {
                R[p] -= FEEDBACK_R_VAR / V[p];
                if(R[p] < 1.0f) R[p] = 1.0f;
            }
            int pc = 0;
            for(int c = 0; c < CH; ++c) pc += popcount16(intra[c]);
            if(pc >= (CH == 1 ? 2 : 4)) ++nonzero_desc;
            for(int c = 0; c < CH; ++c) { ld[c] = intra[c]; last_color[p * CH + c] = cur[c]; }
        }
        for(const NbWrite& w : queue) { // snapshot mode: deferred neighbour writes, raster order of the source
            for(int c = 0; c < CH; ++c) { bgd(w.slot)[w.target * CH + c] = w.desc[c]; bgc(w.slot)[w.target * CH + c] = w.col[c]; }
            ++stats.sample_writes;
        }
        stats.roi_px += roi_count; ++stats.frames;
        postprocess(img, fgmask, aLT, aST, nonzero_desc);
    }

    /// SuBSENSE.cpp:536-611
    void postprocess(const uchar* img, uchar* fgmask, float aLT, float aST, size_t nonzero_desc) {
        std::vector<uchar> cur_blink(npx), preflood(npx), flooded(npx), tmp(npx), cur(raw_mask), dil(npx);
        for(size_t i = 0; i < npx; ++i) {
            cur_blink[i] = raw_mask[i] ^ last_raw[i];
            blinks[i] = cur_blink[i] | last_raw_blink[i];
        }
        last_raw_blink = cur_blink;
        last_raw = raw_mask;
        morph_rect(raw_mask.data(), tmp.data(), W, H, 1, true);      // MORPH_CLOSE 3x3
        morph_rect(tmp.data(), preflood.data(), W, H, 1, false);
        flooded = preflood;
        floodfill_from_origin(flooded.data(), W, H);
        for(size_t i = 0; i < npx; ++i) flooded[i] = (uchar)~flooded[i];
        morph_rect(preflood.data(), tmp.data(), W, H, 3, false);    // erode x3
        for(size_t i = 0; i < npx; ++i) cur[i] = raw_mask[i] | flooded[i] | tmp[i];
        median_binary(cur.data(), last_fg.data(), W, H, median_k);
        morph_rect(last_fg.data(), dil.data(), W, H, 3, true);      // dilate x3
        for(size_t i = 0; i < npx; ++i) {
            blinks[i] &= dil_inv[i];
            dil_inv[i] = (uchar)~dil[i];
            blinks[i] &= dil_inv[i];
        }
        std::memcpy(fgmask, last_fg.data(), npx);
        { // cv::addWeighted(f32, alpha, u8, beta, 0, CV_32F): double accumulate, one rounding (Appendix E)
            const double a1 = (double)(1.0f - aLT), b1 = (1.0 / 255) * (double)aLT;
            const double a2 = (double)(1.0f - aST), b2 = (1.0 / 255) * (double)aST;
            for(size_t i = 0; i < npx; ++i) {
                finLT[i] = (float)((double)finLT[i] * a1 + (double)last_fg[i] * b1);
                finST[i] = (float)((double)finST[i] * a2 + (double)last_fg[i] * b2);
            }
        }
        const float ratio = (float)nonzero_desc / roi_count;
        const size_t lbspOff = (size_t)P.lbsp_threshold_offset;
        if(ratio < LBSPDESC_NONZERO_RATIO_MIN && last_nonzero_ratio < LBSPDESC_NONZERO_RATIO_MIN) {
            for(size_t t = 0; t < 256; ++t)
                if(lut[t] > sat_u8((float)lbspOff + std::ceil((float)t * P.rel_lbsp_threshold / 4))) --lut[t];
        } else if(ratio > LBSPDESC_NONZERO_RATIO_MAX && last_nonzero_ratio > LBSPDESC_NONZERO_RATIO_MAX) {
            for(size_t t = 0; t < 256; ++t)
                if(lut[t] < sat_u8((float)lbspOff + 255 * P.rel_lbsp_threshold)) ++lut[t];
        }
        last_nonzero_ratio = ratio;
        if(lr_scaling) {
            resize_area(img);
            // cv::accumulateWeighted (u8 -> f32): dst = src*a + dst*(1-a), float, no fused multiply-add
            const float bLT = 1.0f - aLT, bST = 1.0f - aST;
            for(size_t i = 0; i < dsLT.size(); ++i) {
                const float sLT = (float)ds_frame[i] * aLT, dLT = dsLT[i] * bLT;
                dsLT[i] = sLT + dLT;
                const float sST = (float)ds_frame[i] * aST, dST = dsST[i] * bST;
                dsST[i] = sST + dST;
            }
            size_t tot_diff = 0;
            for(int i = 0; i < dsW * dsH; ++i) {
                if(C == 1) tot_diff += (size_t)std::fabs(dsST[i] - dsLT[i]) / 2;
                else {
                    const size_t d0 = (size_t)std::fabs(dsST[i * 3] - dsLT[i * 3]);
                    const size_t d1 = (size_t)std::fabs(dsST[i * 3 + 1] - dsLT[i * 3 + 1]);
                    const size_t d2 = (size_t)std::fabs(dsST[i * 3 + 2] - dsLT[i * 3 + 2]);
                    tot_diff += std::max(d0, std::max(d1, d2));
                }
            }
            const float diff_ratio = (float)tot_diff / (dsH * dsW);
            const size_t fl_thr = (size_t)P.color_dist_threshold / 2;
            if(auto_reset) {
                if(frames_since_reset > 1000) auto_reset = false;
                else if(diff_ratio >= fl_thr && reset_cooldown == 0) {
                    frames_since_reset = 0;
                    refresh_model(0.1f, false);
                    reset_cooldown = (size_t)P.n_samples_for_moving_avgs / 4;
                    std::fill(T.begin(), T.end(), 1.0f);
                } else ++frames_since_reset;
            } else if(diff_ratio >= fl_thr * 2) {
                frames_since_reset = 0;
                auto_reset = true;
            }
            if(diff_ratio >= fl_thr / 2) {
                // the reference shifts an int by (int)(ratio/2), undefined for >=32 (ratio >= 64); we define it as 0
                const int sh = (int)(diff_ratio / 2);
                t_lower = (float)std::max(sh < 31 ? ((int)FEEDBACK_T_LOWER >> sh) : 0, 1);
                t_upper = (float)std::max(sh < 31 ? ((int)FEEDBACK_T_UPPER >> sh) : 0, 1);
            } else { t_lower = FEEDBACK_T_LOWER; t_upper = FEEDBACK_T_UPPER; }
            if(reset_cooldown > 0) --reset_cooldown;
        }
    }

    void resize_area(const uchar* img) {
        if(W % 8 == 0 && H % 8 == 0) resize_area_exact(img, W, H, C, 8, ds_frame.data());   // OpenCV's integer-scale fast path
        else resize_area_general(img, W, H, C, dsW, dsH, ds_frame.data());
    }

    void apply(const uchar* img, uchar* fgmask, double lr) {
        if(!initialized) throw std::runtime_error("algo & model must be initialized first");
        if(C == 1) apply_impl<1>(img, fgmask, lr); else apply_impl<3>(img, fgmask, lr);
    }

    /// SuBSENSE.cpp:614-630 (float mean, convertTo round-half-even + saturate)
    void get_background_image(uchar* out) const {
        std::vector<float> acc(npx * C, 0.f);
        for(int s = 0; s < P.n_samples; ++s) {
            const uchar* b = bg_color.data() + (size_t)s * npx * C;
            for(size_t i = 0; i < npx * C; ++i) acc[i] += ((float)b[i]) / P.n_samples;
        }
        for(size_t i = 0; i < npx * C; ++i) out[i] = sat_u8(acc[i]);
    }
    /// SuBSENSE.cpp:632-649
    void get_background_desc_image(ushort* out) const {
        std::vector<float> acc(npx * C, 0.f);
        for(int s = 0; s < P.n_samples; ++s) {
            const ushort* b = bg_desc.data() + (size_t)s * npx * C;
            for(size_t i = 0; i < npx * C; ++i) acc[i] += ((float)b[i]) / P.n_samples;
        }
        for(size_t i = 0; i < npx * C; ++i) { const long r = std::lrint((double)acc[i]); out[i] = (ushort)(r < 0 ? 0 : (r > 65535 ? 65535 : r)); }
    }
};

// =================================================================================================
// LOBSTER (CPU path, video/src/BackgroundSubtractorLOBSTER.cpp:410-620)
// =================================================================================================
struct LOBSTER : BgsBase {
    std::vector<uchar> bg_color;
    std::vector<ushort> bg_desc;
    std::vector<uchar> raw_mask;
    uchar* bgc(int s) { return bg_color.data() + (size_t)s * npx * C; }
    ushort* bgd(int s) { return bg_desc.data() + (size_t)s * npx * C; }

    /// LOBSTER.cpp:410-441 : like SuBSENSE's but the sampled pixel's descriptor is recomputed from
    /// m_oLastColorFrame at refresh time (:428-433) and written back to m_oLastDescFrame
    void refresh_model(float frac, bool force_fg) {
        if(!initialized) throw std::runtime_error("algo must be initialized first");
        if(!(frac > 0.0f && frac <= 1.0f)) throw std::runtime_error("model refresh must be given as a non-null fraction");
        const size_t N = (size_t)P.n_samples;
        const size_t n_refresh = frac < 1.0f ? (size_t)(frac * N) : N;
        const uint32_t epoch = refresh_epoch++;
        size_t start = 0;
        if(frac < 1.0f) start = (size_t)(mode == MODE_REFERENCE ? draw_ref() : philox_draw(seed, epoch, 0, 0, DOM_REFRESH_START)) % N;
        for(size_t p = 0; p < npx; ++p) {
            if(!roi[p]) continue;
            if(force_fg || !last_fg[p]) {
                const int ox = (int)(p % W), oy = (int)(p / W);
                for(size_t s = start; s < start + n_refresh; ++s) {
                    const size_t rs = s % N;
                    const int rnd = mode == MODE_REFERENCE ? draw_ref() : philox_draw(seed, epoch, (uint32_t)p, (uint32_t)rs, DOM_REFRESH);
                    int sx, sy;
                    sample_pos_7x7(rnd, sx, sy, ox, oy, 2, W, H);
                    const size_t sp = (size_t)sy * W + sx;
                    if(force_fg || !last_fg[sp]) {
                        for(int c = 0; c < C; ++c) {
                            const uchar col = last_color[sp * C + c];
                            bgc((int)rs)[p * C + c] = col;
                            uchar vals[16];
                            lbsp_lookup(last_color.data(), W, C, sx, sy, c, vals);
                            last_desc[sp * C + c] = lbsp_threshold(vals, col, lut[col]);
                            bgd((int)rs)[p * C + c] = last_desc[sp * C + c];
                        }
                    }
                }
            }
        }
    }

    /// LOBSTER.cpp:443-457
    void initialize(const uchar* img, int w, int h, int c, const uchar* roi_or_null) {
        initialize_common(img, w, h, c, roi_or_null);
        bg_color.assign((size_t)P.n_samples * npx * C, 0);
        bg_desc.assign((size_t)P.n_samples * npx * C, 0);
        raw_mask.assign(npx, 0);
        initialized = true;
        refresh_model(1.0f, true);
    }

    struct NbWrite { size_t target; int slot; uchar col[3]; ushort desc[3]; };

    /// LOBSTER.cpp:459-581
    template<int CH> void apply_impl(const uchar* img, uchar* fgmask, double lr) {
        if(!(lr > 0)) throw std::runtime_error("learning rate must be a positive value; faster learning is achieved with smaller values");
        const size_t N = (size_t)P.n_samples, REQ = (size_t)P.n_required;
        const size_t LR = std::isinf(lr) ? SIZE_MAX : (size_t)std::ceil(lr);
        const size_t colorThr = (size_t)P.color_dist_threshold, descThr = (size_t)P.desc_dist_threshold;
        const size_t totD = descThr * 3, totC = colorThr * 3, scD = totD / 2, scC = totC / 2;
        std::fill(raw_mask.begin(), raw_mask.end(), 0);
        ++frame_idx; // not in the reference (LOBSTER never touches m_nFrameIdx); used only as the Philox counter
        const uint32_t fr = (uint32_t)frame_idx;
        std::vector<NbWrite> queue;
        for(size_t p = 0; p < npx; ++p) {
            if(!roi[p]) continue;
            const int x = (int)(p % W), y = (int)(p / W);
            const uchar* cur = img + p * CH;
            auto draw = [&](uint32_t site) -> size_t {
                return (size_t)(mode == MODE_REFERENCE ? draw_ref() : philox_draw(seed, fr, (uint32_t)p, site, DOM_APPLY));
            };
            uchar vals[CH][16];
            for(int c = 0; c < CH; ++c) lbsp_lookup(img, W, CH, x, y, c, vals[c]);
            size_t good = 0, s = 0;
            while(good < REQ && s < N) {
                const uchar* bc = bgc((int)s) + p * CH;
                const ushort* bd = bgd((int)s) + p * CH;
                bool ok = true;
                if(CH == 1) {
                    if((size_t)L1dist_u8(cur[0], bc[0]) > colorThr / 2) ok = false;
                    else if((size_t)hdist16(lbsp_threshold(vals[0], bc[0], lut[bc[0]]), bd[0]) > descThr) ok = false;
                } else {
                    size_t tc = 0, td = 0;
                    for(int c = 0; c < CH; ++c) {
                        const size_t cd = L1dist_u8(cur[c], bc[c]);
                        if(cd > scC) { ok = false; break; }
                        const size_t dd = (size_t)hdist16(lbsp_threshold(vals[c], bc[c], lut[bc[c]]), bd[c]);
                        if(dd > scD) { ok = false; break; }
                        tc += cd; td += dd;
                    }
                    if(ok && !(td <= totD && tc <= totC)) ok = false;
                }
                if(ok) ++good;
                ++s;
            }
            stats.samples_scanned += s;
            if(good < REQ) { raw_mask[p] = 255; ++stats.fg_px; }
            else {
                ushort intra[CH];
                for(int c = 0; c < CH; ++c) intra[c] = lbsp_threshold(vals[c], cur[c], lut[cur[c]]);
                if((draw(0) % LR) == 0) {
                    const size_t slot = draw(1) % N;
                    for(int c = 0; c < CH; ++c) { bgc((int)slot)[p * CH + c] = cur[c]; bgd((int)slot)[p * CH + c] = intra[c]; }
                    ++stats.sample_writes;
                }
                if((draw(2) % LR) == 0) {
                    int nx, ny;
                    neighbor_pos_3x3((int)draw(3), nx, ny, x, y, 2, W, H);
                    const size_t slot = mode == MODE_REFERENCE ? draw(4) % N : (draw(1) / N) % N; // snapshot: one Philox block per pixel
                    const size_t q = (size_t)ny * W + nx;
                    if(mode == MODE_REFERENCE) {
                        for(int c = 0; c < CH; ++c) { bgc((int)slot)[q * CH + c] = cur[c]; bgd((int)slot)[q * CH + c] = intra[c]; }
                        ++stats.sample_writes;
                    } else {
                        NbWrite w; w.target = q; w.slot = (int)slot;
                        for(int c = 0; c < CH; ++c) { w.col[c] = cur[c]; w.desc[c] = intra[c]; }
                        queue.push_back(w);
                    }
                }
            }
        }
        for(const NbWrite& w : queue) {
            for(int c = 0; c < CH; ++c) { bgd(w.slot)[w.target * CH + c] = w.desc[c]; bgc(w.slot)[w.target * CH + c] = w.col[c]; }
            ++stats.sample_writes;
        }
        stats.roi_px += roi_count; ++stats.frames;
        median_binary(raw_mask.data(), last_fg.data(), W, H, P.median_blur_kernel_size);
        std::memcpy(fgmask, last_fg.data(), npx);
        std::memcpy(last_color.data(), img, npx * CH); // whole frame, not only the ROI (:580)
    }

    void apply(const uchar* img, uchar* fgmask, double lr) {
        if(!initialized) throw std::runtime_error("algo & model must be initialized first");
        if(C == 1) apply_impl<1>(img, fgmask, lr); else apply_impl<3>(img, fgmask, lr);
    }
    /// LOBSTER.cpp:583-600
    void get_background_image(uchar* out) const {
        std::vector<float> acc(npx * C, 0.f);
        for(int s = 0; s < P.n_samples; ++s) {
            const uchar* b = bg_color.data() + (size_t)s * npx * C;
            for(size_t i = 0; i < npx * C; ++i) acc[i] += ((float)b[i]) / P.n_samples;
        }
        for(size_t i = 0; i < npx * C; ++i) out[i] = sat_u8(acc[i]);
    }
    /// LOBSTER.cpp:602-620
    void get_background_desc_image(ushort* out) const {
        std::vector<float> acc(npx * C, 0.f);
        for(int s = 0; s < P.n_samples; ++s) {
            const ushort* b = bg_desc.data() + (size_t)s * npx * C;
            for(size_t i = 0; i < npx * C; ++i) acc[i] += ((float)b[i]) / P.n_samples;
        }
        for(size_t i = 0; i < npx * C; ++i) { const long r = std::lrint((double)acc[i]); out[i] = (ushort)(r < 0 ? 0 : (r > 65535 ? 65535 : r)); }
    }
};

} // namespace lvo
