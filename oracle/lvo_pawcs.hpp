// ORACLE — TEST INFRASTRUCTURE ONLY (see lvo_common.hpp header).
//
// CPU restatement of BackgroundSubtractorPAWCS (reference: modules/video/src/BackgroundSubtractorPAWCS.cpp, cited as
// PAWCS.cpp:line below; header modules/video/include/litiv/video/BackgroundSubtractorPAWCS.hpp).
// Parity pinned: MODE_REFERENCE equals the reference's own BackgroundSubtractorPAWCS.cpp (oracle/_ref) bit for bit, dictionaries included
// (tests/test_ref_pin_cpu.py); the reference holds no test or golden vector for PAWCS.
//
// Two modes, like the SuBSENSE oracle:
//   MODE_REFERENCE : the reference's raster order, libc rand() clone, every in-loop coupling (neighbour dictionaries,
//                    illumination mask, global dictionary) exactly as written.
//   MODE_SNAPSHOT  : the deterministic parallel semantics the GPU implements:
//     * every pixel reads frame-start state of everything it does not own; own-dictionary updates apply at once;
//     * Philox draws: domain DOM_PAWCS_A, sites 0 (global-word draw), 1 (replace draw), 2 (neighbour draw), 3 (neighbour
//       position), 4+i (illumination update of local word i); domain DOM_PAWCS_B site j (neighbour word j, keyed by the
//       SOURCE pixel);
//     * illumination mask: new[p] = did[p+1] ? (1 & roi[p]) : did[p]  (the reference's write to p-1/p+1 seen by later
//       pixels of the same frame is dropped);
//     * global dictionary: matches/rescues read the frame-start dictionary; occupancy updates are applied after the pixel
//       pass in raster order (per map cell), word weights accumulate in 2^-32 fixed point (order independent); if several
//       pixels ask to replace the last global word, the first one in raster order wins and the others are dropped;
//     * neighbour-dictionary updates are queued and applied after the pixel pass in raster order of their SOURCE pixel;
//       they see the target's current-frame raw mask and intra descriptor;
//     * frame-level float sums (motion analysis, cv::sum of an occupancy map) are fixed-point sums (order independent).
#pragma once
#include "lvo_subsense.hpp"
#include <cfloat>
#include <climits>

namespace lvo {

struct PAWCS : BgsBase {
    // PAWCS.cpp:26-74
    static constexpr float FEEDBACK_R_VAR = 0.01f, FEEDBACK_V_INCR = 1.0f, FEEDBACK_V_DECR = 0.1f;
    static constexpr float FEEDBACK_T_DECR = 0.25f, FEEDBACK_T_INCR = 0.5f, FEEDBACK_T_LOWER = 1.0f, FEEDBACK_T_UPPER = 256.0f;
    static constexpr float UNSTABLE_REG_RATIO_MIN = 0.1f, UNSTABLE_REG_RDIST_MIN = 3.0f;
    static constexpr float LBSPDESC_RATIO_MIN = 0.1f, LBSPDESC_RATIO_MAX = 0.5f;
    static constexpr int FRAMELEVEL_MIN_L1DIST_THRES = 45, FRAMELEVEL_MIN_CDIST_THRES = 45 / 10;
    static constexpr size_t BOOTSTRAP_WIN = 500, RESAMPLING_RATE = 16, LWORD_WEIGHT_OFFSET = BOOTSTRAP_WIN * 2;

    int NW = 0, NG = 0;            // current local / global word counts
    size_t weight_offset = LWORD_WEIGHT_OFFSET;
    bool moving_camera = false;
    float last_nonflat_ratio = 0.0f;
    int median_k = 9, dsW = 0, dsH = 0, gW = 0, gH = 0;
    std::vector<uchar> ds_roi; size_t ds_roi_count = 0;
    std::vector<float> T, R, V, DminLT, DminST, rawLT, rawST, finLT, finST, dsLT, dsST;
    std::vector<uchar> unstable, illum, blinks, last_raw, last_raw_blink, dil, dil_inv, raw_mask, ds_frame;
    // local dictionaries, [p*NW + i] in dictionary order (the reference's pointer array dereferenced)
    std::vector<uint32_t> lw_first, lw_last, lw_occ;
    std::vector<uchar> lw_color, lw_valid;  // [..*C]
    std::vector<ushort> lw_desc;
    // global dictionary: storage indexed by word identity; gdict = dictionary order -> identity
    std::vector<float> gw_weight, gw_map;   // map: [g][gH*gW]
    std::vector<uchar> gw_bits, gw_color, gw_valid;
    std::vector<ushort> gw_desc;
    std::vector<int32_t> gdict;
    std::vector<uchar> glut;                // [p*NG + i] identity, per-pixel sort LUT
    int g_created = 0;

    float lweight(size_t k, size_t frame) const { // PAWCS.cpp:1596-1598
        return (float)lw_occ[k] / (float)(((size_t)lw_last[k] - lw_first[k]) + (frame - lw_last[k]) * 2 + weight_offset);
    }
    void lswap(size_t a, size_t b) {
        std::swap(lw_first[a], lw_first[b]); std::swap(lw_last[a], lw_last[b]); std::swap(lw_occ[a], lw_occ[b]); std::swap(lw_valid[a], lw_valid[b]);
        for(int c = 0; c < C; ++c) { std::swap(lw_color[a * C + c], lw_color[b * C + c]); std::swap(lw_desc[a * C + c], lw_desc[b * C + c]); }
    }
    size_t cell_of(size_t p) const { return (size_t)((p / W) / 2) * gW + (size_t)((p % W) / 2); } // PAWCS.cpp:525
    float* gmap(int g) { return gw_map.data() + (size_t)g * gW * gH; }

    template<int CH> static size_t color_dist(const uchar* a, const uchar* b, size_t& l1, size_t& cd) {
        if(CH == 1) { l1 = L1dist_u8(a[0], b[0]); cd = 0; return l1; }
        l1 = L1dist_arr_u8<CH>(a, b); cd = cdist_u8<(CH == 1 ? 3 : CH)>(a, b); return cmixdist(l1, cd); // Q1: wrapping uchar L1
    }
    template<int CH> static size_t desc_hdist(const ushort* a, const ushort* b) { size_t r = 0; for(int c = 0; c < CH; ++c) r += (size_t)hdist16(a[c], b[c]); return r; }
    template<int CH> static size_t desc_bits(const ushort* a) { size_t r = 0; for(int c = 0; c < CH; ++c) r += (size_t)popcount16(a[c]); return r; }
    static size_t flat_bits(int CH) { return CH == 1 ? 2 : 4; } // FLAT_REGION_BIT_COUNT (16/8) [*2 for 3ch]

    template<int CH> void thresholds(size_t p, size_t& thrC, size_t& thrD) const { // PAWCS.cpp:664-665 / :993-994
        // `sqrt(float)*size_t`: with OpenCV's headers in scope (<math.h> C++ wrapper) the float overload is selected and the
        // product is a float (assumption; a double sqrt could differ only when the product sits within 1 ulp of an integer)
        const float f = std::sqrt(R[p]) * (float)(size_t)P.color_dist_threshold;
        const size_t base = ((size_t)1 << ((size_t)std::floor(R[p] + 0.5f))) + (size_t)P.desc_dist_threshold + (unstable[p] ? (size_t)P.desc_dist_threshold : 0);
        if(CH == 1) { thrC = (size_t)f / 2; thrD = base; } else { thrC = (size_t)f * 3; thrD = base * 3; }
    }

    // ------------------------------------------------------------------------------------------
    // refreshModel (PAWCS.cpp:107-429)
    // ------------------------------------------------------------------------------------------
    template<int CH> void refresh_impl(size_t base_occ, float decr_frac, bool force) {
        const uint32_t epoch = refresh_epoch++;
        const uint32_t fr = (uint32_t)frame_idx;
        for(size_t p = 0; p < npx; ++p) {
            if(!roi[p]) continue;
            if(!(force || !dil[p])) continue;
            uint32_t site = 0;
            auto draw = [&]() -> size_t { return (size_t)(mode == MODE_REFERENCE ? draw_ref() : philox_draw(seed, epoch, (uint32_t)p, site++, DOM_REFRESH)); };
            const size_t ld = p * NW;
            size_t thrC, thrD; thresholds<CH>(p, thrC, thrD);
            if(decr_frac > 0.0f) // :116-122 / :271-277
                for(int i = 0; i < NW; ++i) if(lw_valid[ld + i]) lw_occ[ld + i] -= (uint32_t)(size_t)(decr_frac * lw_occ[ld + i]);
            const int ox = (int)(p % W), oy = (int)(p / W);
            for(int it = 0; it < 7 * 7 * 2; ++it) { // :125-163 / :280-320
                int sx, sy;
                sample_pos_7x7((int)draw(), sx, sy, ox, oy, 2, W, H);
                const size_t sp = (size_t)sy * W + sx;
                if(!(force || !dil[sp])) continue;
                const uchar* scol = last_color.data() + sp * CH; const ushort* sdesc = last_desc.data() + sp * CH;
                bool found_uninit = false; int i;
                for(i = 0; i < NW; ++i) {
                    const size_t k = ld + i;
                    size_t l1, cd;
                    if(lw_valid[k] && color_dist<CH>(scol, lw_color.data() + k * CH, l1, cd) <= thrC && desc_hdist<CH>(sdesc, lw_desc.data() + k * CH) <= thrD) {
                        lw_occ[k] += 1; lw_last[k] = fr; break;
                    } else if(!lw_valid[k]) found_uninit = true;
                }
                if(i == NW) {
                    i = NW - 1; (void)found_uninit; // a new list entry and the in-place overwrite are the same slot in this layout
                    const size_t k = ld + i;
                    for(int c = 0; c < CH; ++c) { lw_color[k * CH + c] = scol[c]; lw_desc[k * CH + c] = sdesc[c]; }
                    lw_occ[k] = (uint32_t)base_occ; lw_first[k] = fr; lw_last[k] = fr; lw_valid[k] = 1;
                }
                while(i > 0 && (!lw_valid[ld + i - 1] || lweight(ld + i, fr) > lweight(ld + i - 1, fr))) { lswap(ld + i, ld + i - 1); --i; }
            }
            for(int i = 1; i < NW; ++i) { // :165-180 / :322-339 random resampling of the words still missing
                const size_t k = ld + i;
                if(lw_valid[k]) continue;
                const size_t r = draw() % (size_t)i;
                const size_t kr = ld + r;
                const int off = CH == 1 ? (int)(draw() % (thrC + 1)) - (int)thrC / 2 : (int)(draw() % (thrC / 3 + 1)) - (int)(thrC / 6);
                for(int c = 0; c < CH; ++c) { lw_color[k * CH + c] = sat_u8_int((long)lw_color[kr * CH + c] + off); lw_desc[k * CH + c] = lw_desc[kr * CH + c]; }
                lw_occ[k] = (uint32_t)std::max((size_t)((float)lw_occ[kr] * ((float)(NW - i) / (float)NW)), (size_t)1);
                lw_first[k] = fr; lw_last[k] = fr; lw_valid[k] = 1;
            }
        }
        // global resampling (:183-256 / :342-408): sequential by nature (<= ~4*NG pixels), identical in both modes
        size_t incr = std::max(npx / (size_t)NG, (size_t)1);
        for(int pass = 0; pass < 2; ++pass) {
            for(size_t p = 0; p < npx; ++p) {
                if(!roi[p] || (p % incr) != 0) continue;
                if(!(force || !dil[p])) continue;
                const size_t ld = p * NW;
                size_t thrC, thrD; thresholds<CH>(p, thrC, thrD);
                const float bw = lweight(ld, fr);
                const uchar bits = (uchar)desc_bits<CH>(lw_desc.data() + ld * CH);
                bool found_uninit = false; int i;
                for(i = 0; i < NG; ++i) {
                    const int g = gdict[i];
                    size_t l1, cd;
                    if(g >= 0 && (size_t)L1dist_u8(bits, gw_bits[g]) <= thrD / 4 && color_dist<CH>(lw_color.data() + ld * CH, gw_color.data() + (size_t)g * CH, l1, cd) <= thrC) break;
                    else if(g < 0) found_uninit = true;
                }
                if(i == NG) {
                    i = NG - 1;
                    int g = found_uninit ? g_created++ : gdict[i];
                    for(int c = 0; c < CH; ++c) { gw_color[(size_t)g * CH + c] = lw_color[ld * CH + c]; gw_desc[(size_t)g * CH + c] = lw_desc[ld * CH + c]; }
                    gw_bits[g] = bits; std::fill(gmap(g), gmap(g) + (size_t)gW * gH, 0.0f); gw_weight[g] = 0.0f; gw_valid[g] = 1;
                    gdict[i] = g;
                }
                const int g = gdict[i];
                float& cellw = gmap(g)[cell_of(p)];
                if(cellw < bw) { gw_weight[g] += bw; cellw += bw; }
                while(i > 0 && (gdict[i - 1] < 0 || gw_weight[gdict[i]] > gw_weight[gdict[i - 1]])) { std::swap(gdict[i], gdict[i - 1]); --i; }
            }
            incr = std::max(incr / 3, (size_t)1);
        }
        for(int i = 0; i < NG; ++i) if(gdict[i] < 0) { // :246-255 / :397-408
            const int g = g_created++;
            for(int c = 0; c < CH; ++c) { gw_color[(size_t)g * CH + c] = 0; gw_desc[(size_t)g * CH + c] = 0; }
            gw_bits[g] = 0; std::fill(gmap(g), gmap(g) + (size_t)gW * gH, 0.0f); gw_weight[g] = 0.0f; gw_valid[g] = 1; gdict[i] = g;
        }
        glut_bubble_pass(); // :411-428
    }
    void glut_bubble_pass() { // PAWCS.cpp:416-427 / :1322-1333
        for(size_t p = 0; p < npx; ++p) {
            if(!roi[p]) continue;
            uchar* l = glut.data() + p * NG;
            const size_t cell = cell_of(p);
            float last = gmap(l[0])[cell];
            for(int i = 1; i < NG; ++i) {
                const float w = gmap(l[i])[cell];
                if(w > last) std::swap(l[i], l[i - 1]); else last = w;
            }
        }
    }
    void refresh_model(size_t base_occ, float decr_frac, bool force) {
        if(!initialized) throw std::runtime_error("algo must be initialized first");
        if(!(decr_frac >= 0.0f && decr_frac <= 1.0f)) throw std::runtime_error("model occurrence decrementation must be given as a non-null fraction");
        if(C == 1) refresh_impl<1>(base_occ, decr_frac, force); else refresh_impl<3>(base_occ, decr_frac, force);
    }

    /// cv::resize(src, dst, Size(W/8, H/8), 0, 0, INTER_AREA): OpenCV's integer-scale fast path when both dimensions divide by 8,
    /// its general area path otherwise (lvo_common.hpp)
    void ds_area(const uchar* src, int ch, uchar* dst) const {
        if(W % 8 == 0 && H % 8 == 0) resize_area_exact(src, W, H, ch, 8, dst);
        else resize_area_general(src, W, H, ch, dsW, dsH, dst);
    }

    // ------------------------------------------------------------------------------------------
    // initialize (PAWCS.cpp:431-557)
    // ------------------------------------------------------------------------------------------
    void initialize(const uchar* img, int w, int h, int c, const uchar* roi_or_null) {
        if(P.n_samples <= 0 || P.n_samples / 2 <= 0) throw std::runtime_error("max local/global word counts must be positive");
        initialize_common(img, w, h, c, roi_or_null);
        moving_camera = false; auto_reset = true;
        dsW = W / 8; dsH = H / 8; gW = W / 2; gH = H / 2;
        ds_roi.assign((size_t)dsW * dsH, 0);
        ds_area(roi.data(), 1, ds_roi.data());
        last_nonflat_ratio = 0.0f;
        NW = P.n_samples;
        const int maxG = P.n_samples / 2, qvga = 320 * 240, defk = P.median_blur_kernel_size;
        if(orig_roi_count >= npx / 2 && (int)npx >= qvga) {
            const float sc = (float)npx / qvga;
            const int rawk = std::min((int)std::floor(0.5f + sc) + defk, defk + 4);
            median_k = (rawk % 2) ? rawk : rawk - 1;
            NG = maxG;
            for(auto& v : ds_roi) v |= 127;
        } else {
            const float sc = (float)orig_roi_count / qvga;
            const int rawk = std::min((int)std::floor(0.5f + defk * sc * 2) + (defk - 4), defk);
            median_k = (rawk % 2) ? rawk : rawk - 1;
            NG = (int)std::min((size_t)std::pow((double)((float)maxG * sc), 2.0) + 1, (size_t)maxG);
        }
        if(median_k < 1) median_k = 1;
        if(C == 1) { NW = std::max(NW / 2, 1); NG = std::max(NG / 2, 1); }
        ds_roi_count = 0; for(uchar v : ds_roi) ds_roi_count += v != 0;
        weight_offset = LWORD_WEIGHT_OFFSET;
        illum.assign(npx, 0);
        T.assign(npx, FEEDBACK_T_LOWER); R.assign(npx, 2.0f); V.assign(npx, FEEDBACK_V_INCR * 10);
        DminLT.assign(npx, 0.f); DminST.assign(npx, 0.f); rawLT.assign(npx, 0.f); rawST.assign(npx, 0.f); finLT.assign(npx, 0.f); finST.assign(npx, 0.f);
        dsLT.assign((size_t)dsW * dsH * C, 0.f); dsST.assign((size_t)dsW * dsH * C, 0.f); ds_frame.assign((size_t)dsW * dsH * C, 0);
        unstable.assign(npx, 0); blinks.assign(npx, 0); last_raw.assign(npx, 0); last_raw_blink.assign(npx, 0);
        dil.assign(npx, 0); dil_inv.assign(npx, 0); raw_mask.assign(npx, 0);
        lw_first.assign(npx * NW, 0); lw_last.assign(npx * NW, 0); lw_occ.assign(npx * NW, 0); lw_valid.assign(npx * NW, 0);
        lw_color.assign(npx * NW * C, 0); lw_desc.assign(npx * NW * C, 0);
        gw_weight.assign(NG, 0.f); gw_map.assign((size_t)NG * gW * gH, 0.f); gw_bits.assign(NG, 0); gw_valid.assign(NG, 0);
        gw_color.assign((size_t)NG * C, 0); gw_desc.assign((size_t)NG * C, 0);
        gdict.assign(NG, -1); g_created = 0;
        glut.assign(npx * NG, 0);
        for(size_t p = 0; p < npx; ++p) for(int i = 0; i < NG; ++i) glut[p * NG + i] = (uchar)i;
        initialized = true;
        refresh_model(1, 0.0f, false);
    }

    struct GOp { uint32_t p; int g; float w; };            // g < 0: replace the last global word
    struct NbIntent { uint32_t src, target; size_t thrC, thrD, rate; float wthr; };

    // neighbour dictionary update (PAWCS.cpp:838-888 / :1176-1246) of target q by source p
    template<int CH> void neighbor_update(size_t p, size_t q, const uchar* cur, const ushort* intra, size_t thrC, size_t thrD, size_t rate, float wthr,
                                          bool flat, size_t occ_incr, bool boot, uint32_t fr) {
        const size_t ld = q * NW;
        const float init_w = 1.0f / (float)weight_offset;
        float sum = 0.0f;
        for(int j = 0; j < NW && sum < wthr; ++j) {
            const size_t k = ld + j;
            size_t l1, cd;
            const size_t mix = color_dist<CH>(cur, lw_color.data() + k * CH, l1, cd);
            const size_t hd = desc_hdist<CH>(intra, lw_desc.data() + k * CH);
            const bool nflat = desc_bits<CH>(lw_desc.data() + k * CH) < flat_bits(CH);
            const size_t incr = nflat ? occ_incr * 2 : occ_incr;
            const bool writeback = CH != 1; // Q8: the 1-channel path works on a by-value copy of the word (PAWCS.cpp:838)
            auto credit = [&]() {
                const float w = lweight(k, fr);
                sum += w;
                if(writeback) { lw_last[k] = fr; if(w < 1.0f) lw_occ[k] += (uint32_t)incr; }
            };
            if(mix <= thrC && hd <= thrD) credit();
            else if(!raw_mask[q] && flat && (boot || ((size_t)(mode == MODE_REFERENCE ? draw_ref() : philox_draw(seed, fr, (uint32_t)p, (uint32_t)j, DOM_PAWCS_B)) % rate) == 0)) {
                const ushort* qd = last_desc.data() + q * CH;
                const size_t lhd = desc_hdist<CH>(intra, qd);
                if(mix <= thrC && lhd <= thrD / 2) {
                    credit();
                    if(writeback) for(int c = 0; c < CH; ++c) lw_desc[k * CH + c] = intra[c];
                } else if(CH != 1) {
                    const bool lflat = desc_bits<CH>(qd) < flat_bits(CH);
                    if(lflat && flat && lhd + hd <= thrD && cd <= thrC / 4) {
                        credit();
                        for(int c = 0; c < CH; ++c) lw_color[k * CH + c] = cur[c];
                    }
                }
            }
        }
        if(sum < init_w) {
            const size_t k = ld + NW - 1;
            for(int c = 0; c < CH; ++c) { lw_color[k * CH + c] = cur[c]; lw_desc[k * CH + c] = intra[c]; }
            lw_occ[k] = (uint32_t)occ_incr; lw_first[k] = fr; lw_last[k] = fr;
        }
    }

    // ------------------------------------------------------------------------------------------
    // apply (PAWCS.cpp:559-1523)
    // ------------------------------------------------------------------------------------------
    template<int CH> void apply_impl(const uchar* img, uchar* fgmask, double lr_override) {
        const size_t colorRange = CH == 1 ? 255 : 765, descRange = CH == 1 ? 16 : 48;
        std::fill(raw_mask.begin(), raw_mask.end(), 0);
        const bool boot = ++frame_idx <= BOOTSTRAP_WIN;
        const size_t nLT = boot ? (size_t)P.n_samples_for_moving_avgs / 2 : (size_t)P.n_samples_for_moving_avgs, nST = nLT / 4;
        const float aLT = 1.0f / std::min(frame_idx, nLT), aST = 1.0f / std::min(frame_idx, nST);
        const size_t grate = boot ? RESAMPLING_RATE / 2 : RESAMPLING_RATE;
        const uint32_t fr = (uint32_t)frame_idx;
        const bool snap = mode == MODE_SNAPSHOT;
        size_t flat_count = 0;
        std::vector<uchar> did;        // snapshot: illumination updates of this frame
        std::vector<GOp> gops; std::vector<NbIntent> intents;
        if(snap) did.assign(npx, 0);
        const float init_w = 1.0f / (float)weight_offset;
        for(size_t p = 0; p < npx; ++p) {
            if(!roi[p]) continue;
            const int x = (int)(p % W), y = (int)(p / W);
            const uchar* cur = img + p * CH;
            const size_t ld = p * NW;
            const size_t cell = cell_of(p);
            auto drawA = [&](uint32_t site) -> size_t { return (size_t)(snap ? philox_draw(seed, fr, (uint32_t)p, site, DOM_PAWCS_A) : draw_ref()); };
            size_t minColor = colorRange, minDesc = descRange;
            const float best_w = lweight(ld, fr);
            const float wthr = best_w / (R[p] * 2);
            const bool border = roi[p] < 255;
            uchar vals[CH][16]; ushort intra[CH];
            for(int c = 0; c < CH; ++c) { lbsp_lookup(img, W, CH, x, y, c, vals[c]); intra[c] = lbsp_threshold(vals[c], cur[c], lut[cur[c]]); }
            const uchar bits = (uchar)desc_bits<CH>(intra);
            const bool flat = bits < flat_bits(CH);
            if(flat) ++flat_count;
            const size_t occ_incr = (1 + reset_cooldown) << (int)(flat || boot);
            const size_t rate = std::isinf(lr_override) ? SIZE_MAX : (lr_override > 0 ? (size_t)std::ceil(lr_override) : flat ? (size_t)std::ceil(T[p] + FEEDBACK_T_LOWER) / 2 : (size_t)std::ceil(T[p]));
            size_t thrC, thrD; thresholds<CH>(p, thrC, thrD);
            uchar illum_cur = illum[p]; // reference mode: live value (may have been set by p-1 in this frame)
            int i = 0; float sum = 0.0f, last_w = FLT_MAX;
            while(i < NW && sum < wthr) { // :672-727 / :1002-1053
                const size_t k = ld + i;
                const float w = lweight(k, fr);
                size_t l1, cd;
                const size_t mix = color_dist<CH>(cur, lw_color.data() + k * CH, l1, cd);
                const size_t ihd = desc_hdist<CH>(intra, lw_desc.data() + k * CH);
                ushort inter[CH];
                for(int c = 0; c < CH; ++c) inter[c] = lbsp_threshold(vals[c], lw_color[k * CH + c], lut[lw_color[k * CH + c]]);
                const size_t ehd = desc_hdist<CH>(inter, lw_desc.data() + k * CH);
                const size_t dd = (ihd + ehd) / 2;
                if((!unstable[p] || flat || border) && mix <= thrC && l1 >= thrC / 2 && ihd <= thrD / 2
                   && (drawA(4 + (uint32_t)i) % (illum_cur ? (rate / 2 + 1) : rate)) == 0) { // illumination update
                    for(int c = 0; c < CH; ++c) { lw_color[k * CH + c] = cur[c]; lw_desc[k * CH + c] = intra[c]; }
                    if(snap) did[p] = 1;
                    else { illum[p - 1] = 1 & roi[p - 1]; illum[p + 1] = 1 & roi[p + 1]; illum[p] = 2; }
                    illum_cur = 2;
                }
                if(dd <= thrD && mix <= thrC) {
                    sum += w;
                    lw_last[k] = fr;
                    if((!last_fg[p] || moving_camera) && w < 1.0f) lw_occ[k] += (uint32_t)occ_incr;
                    minColor = std::min(minColor, mix); minDesc = std::min(minDesc, dd);
                }
                if(w > last_w) { lswap(k, k - 1); ++stats.sample_writes; } else last_w = w;
                ++i;
            }
            stats.samples_scanned += (uint64_t)i;
            while(i < NW) { // :728-739 / :1054-1065 : the bubble pass continues over the rest of the dictionary
                const float w = lweight(ld + i, fr);
                if(w > last_w) { lswap(ld + i, ld + i - 1); ++stats.sample_writes; } else last_w = w; // stats: dictionary swaps
                ++i;
            }
            uchar seg = 0;
            if(sum >= wthr || border) { // background (:741-780 / :1070-1106)
                const float nmin = std::max((float)minColor / colorRange, (float)minDesc / descRange);
                DminLT[p] = DminLT[p] * (1.0f - aLT) + nmin * aLT; DminST[p] = DminST[p] * (1.0f - aST) + nmin * aST;
                rawLT[p] = rawLT[p] * (1.0f - aLT); rawST[p] = rawST[p] * (1.0f - aST);
                if((drawA(0) % rate) == 0) {
                    int gi, g = -1;
                    for(gi = 0; gi < NG; ++gi) {
                        g = glut[p * NG + gi];
                        size_t l1, cd;
                        if((size_t)L1dist_u8(bits, gw_bits[g]) <= thrD / 4 && color_dist<CH>(cur, gw_color.data() + (size_t)g * CH, l1, cd) <= thrC) break;
                    }
                    const bool found = gi != NG;
                    // rate*2 overflows for the "never" rate; x % (2*rate) == x whenever 2*rate > x
                    auto draw_rep = [&]() { const size_t d = drawA(1); return rate > (SIZE_MAX >> 1) ? d : d % (rate * 2); };
                    if(found || draw_rep() == 0) {
                        if(snap) gops.push_back(GOp{(uint32_t)p, found ? g : -1, sum});
                        else {
                            if(!found) {
                                g = gdict[NG - 1];
                                for(int c = 0; c < CH; ++c) { gw_color[(size_t)g * CH + c] = cur[c]; gw_desc[(size_t)g * CH + c] = intra[c]; }
                                gw_bits[g] = bits; std::fill(gmap(g), gmap(g) + (size_t)gW * gH, 0.0f); gw_weight[g] = 0.0f;
                            }
                            float& cw = gmap(g)[cell];
                            if(cw < sum) { gw_weight[g] += sum; cw += sum; }
                        }
                    }
                }
            } else { // foreground (:781-827 / :1107-1155)
                const float nmin = std::max(std::max((float)minColor / colorRange, (float)minDesc / descRange), (wthr - sum) / wthr);
                DminLT[p] = DminLT[p] * (1.0f - aLT) + nmin * aLT; DminST[p] = DminST[p] * (1.0f - aST) + nmin * aST;
                rawLT[p] = rawLT[p] * (1.0f - aLT) + aLT; rawST[p] = rawST[p] * (1.0f - aST) + aST;
                if(flat || (drawA(0) % rate) == 0) {
                    int gi, g = -1;
                    for(gi = 0; gi < NG; ++gi) {
                        g = glut[p * NG + gi];
                        size_t l1, cd;
                        if((size_t)L1dist_u8(bits, gw_bits[g]) <= thrD / 4 && color_dist<CH>(cur, gw_color.data() + (size_t)g * CH, l1, cd) <= thrC) break;
                    }
                    if(gi == NG) seg = 255;
                    else if(sum + gmap(g)[cell] / (flat ? 2 : 4) < wthr) seg = 255;
                } else seg = 255;
                if(sum < init_w) { // new local word over the last one
                    const size_t k = ld + NW - 1;
                    for(int c = 0; c < CH; ++c) { lw_color[k * CH + c] = cur[c]; lw_desc[k * CH + c] = intra[c]; }
                    lw_occ[k] = (uint32_t)occ_incr; lw_first[k] = fr; lw_last[k] = fr;
                }
            }
            raw_mask[p] = seg;
            if(seg) ++stats.fg_px;
            // neighbour update (:829-889 / :1164-1247)
            if((!seg && (drawA(2) % rate) == 0) || border || moving_camera) {
                int nx, ny;
                const int rpos = (int)drawA(3);
                if(flat || border || moving_camera) neighbor_pos_5x5(rpos, nx, ny, x, y, 2, W, H); else neighbor_pos_3x3(rpos, nx, ny, x, y, 2, W, H);
                const size_t q = (size_t)ny * W + nx;
                if(roi[q]) {
                    if(snap) intents.push_back(NbIntent{(uint32_t)p, (uint32_t)q, thrC, thrD, rate, wthr});
                    else {
                        // reference order: the target's descriptor/mask are the live ones (already updated iff q precedes p)
                        neighbor_update<CH>(p, q, cur, intra, thrC, thrD, rate, wthr, flat, occ_incr, boot, fr);
                    }
                }
            }
            if(snap) { /* own illumination value is resolved after the pass */ }
            else if(illum[p]) illum[p] -= 1;
            // feedback (:890-907 / :1248-1265); the unstable flag was read for the thresholds before this rewrite
            unstable[p] = R[p] > UNSTABLE_REG_RDIST_MIN || (rawLT[p] - finLT[p]) > UNSTABLE_REG_RATIO_MIN || (rawST[p] - finST[p]) > UNSTABLE_REG_RATIO_MIN;
            const float dmin = std::min(DminLT[p], DminST[p]), dmax = std::max(DminLT[p], DminST[p]);
            if(last_fg[p] || (dmin < UNSTABLE_REG_RATIO_MIN && seg)) T[p] = std::min(T[p] + FEEDBACK_T_INCR / (dmax * V[p]), FEEDBACK_T_UPPER);
            else T[p] = std::max(T[p] - FEEDBACK_T_DECR * V[p] / dmax, FEEDBACK_T_LOWER);
            if(dmax > UNSTABLE_REG_RATIO_MIN && blinks[p]) V[p] += boot ? FEEDBACK_V_INCR * 2 : FEEDBACK_V_INCR;
            else V[p] = std::max(V[p] - FEEDBACK_V_DECR * ((boot || flat) ? 2.0f : last_fg[p] ? 0.5f : 1.0f), FEEDBACK_V_DECR);
            if((double)R[p] < std::pow((double)(1.0f + dmin * 2), 2.0)) R[p] += FEEDBACK_R_VAR * (V[p] - FEEDBACK_V_DECR);
            else R[p] = std::max(R[p] - FEEDBACK_R_VAR / V[p], 1.0f);
            for(int c = 0; c < CH; ++c) { last_desc[p * CH + c] = intra[c]; last_color[p * CH + c] = cur[c]; }
        }
        if(snap) {
            // illumination mask of the next frame
            for(size_t p = 0; p < npx; ++p) {
                if(!roi[p]) { illum[p] = 0; continue; }
                illum[p] = (p + 1 < npx && did[p + 1]) ? (uchar)(1 & roi[p]) : (did[p] ? 1 : 0);
            }
            // global dictionary: replacement (first requester wins), then occupancy updates in raster order
            int rep = -1;
            for(size_t n = 0; n < gops.size(); ++n) if(gops[n].g < 0) { rep = (int)n; break; }
            if(rep >= 0) {
                const int g = gdict[NG - 1];
                const size_t p = gops[rep].p;
                const ushort* intra = last_desc.data() + p * CH; // == this frame's intra descriptor
                for(int c = 0; c < CH; ++c) { gw_color[(size_t)g * CH + c] = img[p * CH + c]; gw_desc[(size_t)g * CH + c] = intra[c]; }
                gw_bits[g] = (uchar)desc_bits<CH>(intra); std::fill(gmap(g), gmap(g) + (size_t)gW * gH, 0.0f); gw_weight[g] = 0.0f;
                gops[rep].g = g;
            }
            std::vector<int64_t> acc(NG, 0);
            for(const GOp& o : gops) {
                if(o.g < 0) continue;
                float& cw = gmap(o.g)[cell_of(o.p)];
                if(cw < o.w) { acc[o.g] += (int64_t)std::llrint((double)o.w * 4294967296.0); cw += o.w; }
            }
            for(int g = 0; g < NG; ++g) if(acc[g]) gw_weight[g] = (float)((double)gw_weight[g] + (double)acc[g] / 4294967296.0);
            // queued neighbour-dictionary updates, raster order of the source
            for(const NbIntent& it : intents) {
                const size_t p = it.src;
                const ushort* intra = last_desc.data() + p * CH;
                const bool flat = desc_bits<CH>(intra) < flat_bits(CH);
                const size_t occ_incr = (1 + reset_cooldown) << (int)(flat || boot);
                neighbor_update<CH>(p, it.target, img + p * CH, intra, it.thrC, it.thrD, it.rate, it.wthr, flat, occ_incr, boot, fr);
            }
        }
        stats.roi_px += roi_count; ++stats.frames;
        global_maintenance(grate);
        postprocess<CH>(img, fgmask, aLT, aST, flat_count, boot, nST);
    }

    /// cv::sum of an occupancy map: reference mode = sequential double accumulation; snapshot = 2^-32 fixed point
    float map_sum(int g) {
        const float* m = gmap(g); const size_t n = (size_t)gW * gH;
        if(mode == MODE_REFERENCE) { double s = 0; for(size_t i = 0; i < n; ++i) s += (double)m[i]; return (float)s; }
        int64_t s = 0; for(size_t i = 0; i < n; ++i) s += (int64_t)std::llrint((double)m[i] * 4294967296.0);
        return (float)((double)s / 4294967296.0);
    }
    /// cv::blur(32F, 3x3, BORDER_REPLICATE): row sums and column sums in double, scaled by 1/9, rounded once to float
    void blur3(float* m) const {
        std::vector<double> rs((size_t)gW * gH);
        for(int y = 0; y < gH; ++y) for(int x = 0; x < gW; ++x) {
            const float* r = m + (size_t)y * gW;
            rs[(size_t)y * gW + x] = (double)r[std::max(x - 1, 0)] + (double)r[x] + (double)r[std::min(x + 1, gW - 1)];
        }
        for(int y = 0; y < gH; ++y) for(int x = 0; x < gW; ++x) {
            const double s = rs[(size_t)std::max(y - 1, 0) * gW + x] + rs[(size_t)y * gW + x] + rs[(size_t)std::min(y + 1, gH - 1) * gW + x];
            m[(size_t)y * gW + x] = (float)(s * (1.0 / 9.0));
        }
    }
    /// PAWCS.cpp:1300-1334
    void global_maintenance(size_t grate) {
        const bool recalc = !(frame_idx % (grate << 5)), update = !(frame_idx % grate);
        for(int i = 0; i < NG; ++i) {
            const int g = gdict[i];
            if(recalc && gw_weight[g] > 0.0f) {
                gw_weight[g] = map_sum(g);
                if(gw_weight[g] < 1.0f) { gw_weight[g] = 0.0f; std::fill(gmap(g), gmap(g) + (size_t)gW * gH, 0.0f); }
            }
            if(update && gw_weight[g] > 0.0f) {
                float* m = gmap(g);
                for(int y = 0; y < gH; ++y) for(int x = 0; x < gW; ++x) // accumulateProduct(map, -0.1, map, mask = nearest-downscaled dil_inv)
                    if(dil_inv[(size_t)(y * 2) * W + (x * 2)]) { const float t = m[(size_t)y * gW + x] * -0.1f; m[(size_t)y * gW + x] += t; }
                gw_weight[g] *= 0.9f;
                blur3(m);
            }
            if(i > 0 && gw_weight[gdict[i]] > gw_weight[gdict[i - 1]]) std::swap(gdict[i], gdict[i - 1]);
        }
        if(update) glut_bubble_pass();
    }

    /// masked float distances of the frame-level analysis (math.hpp:257-270 L1dist, :496-527/:545-559 cdist on float arrays)
    float masked_l1(const float* a, const float* b, const uchar* m, bool only255) const {
        const size_t n = (size_t)dsW * dsH;
        float facc = 0.0f; int64_t iacc = 0;
        for(size_t i = 0; i < n; ++i) {
            if(only255 ? m[i] != 255 : m[i] == 0) continue;
            float t = 0.0f;
            for(int c = 0; c < C; ++c) t += std::fabs(a[i * C + c] - b[i * C + c]);
            if(mode == MODE_REFERENCE) facc += t; else iacc += (int64_t)std::llrint((double)t * 65536.0);
        }
        return mode == MODE_REFERENCE ? facc : (float)((double)iacc / 65536.0);
    }
    float masked_cdist3(const float* a, const float* b, const uchar* m) const {
        const size_t n = (size_t)dsW * dsH;
        float facc = 0.0f; int64_t iacc = 0;
        for(size_t i = 0; i < n; ++i) {
            if(m[i] != 255) continue;
            const float* cu = a + i * 3; const float* bg = b + i * 3;
            bool nonconst = false, nonnull = cu[0] != bg[0];
            for(int c = 1; c < 3; ++c) { nonconst |= (cu[c] != cu[c - 1]) || (bg[c] != bg[c - 1]); nonnull |= cu[c] != bg[c]; }
            float t = 0.0f;
            if(nonconst && nonnull) {
                float cs = 0, bs = 0, mix = 0;
                for(int c = 0; c < 3; ++c) { cs += cu[c] * cu[c]; bs += bg[c] * bg[c]; mix += cu[c] * bg[c]; }
                bs += FLT_EPSILON;
                const float q = (mix * mix) / bs;
                if(!(cs <= q)) t = std::sqrt(cs - q);
            }
            if(mode == MODE_REFERENCE) facc += t; else iacc += (int64_t)std::llrint((double)t * 65536.0);
        }
        return mode == MODE_REFERENCE ? facc : (float)((double)iacc / 65536.0);
    }

    /// PAWCS.cpp:1443-1516
    template<int CH> void postprocess(const uchar* img, uchar* fgmask, float aLT, float aST, size_t flat_count, bool boot, size_t nST) {
        std::vector<uchar> cur_blink(npx), preflood(npx), flooded(npx), tmp(npx), cur(raw_mask);
        for(size_t i = 0; i < npx; ++i) { cur_blink[i] = raw_mask[i] ^ last_raw[i]; blinks[i] = cur_blink[i] | last_raw_blink[i]; }
        last_raw_blink = cur_blink; last_raw = raw_mask;
        morph_rect(raw_mask.data(), tmp.data(), W, H, 1, true); morph_rect(tmp.data(), preflood.data(), W, H, 1, false);
        flooded = preflood; floodfill_from_origin(flooded.data(), W, H);
        for(size_t i = 0; i < npx; ++i) flooded[i] = (uchar)~flooded[i];
        morph_rect(preflood.data(), tmp.data(), W, H, 3, false);
        for(size_t i = 0; i < npx; ++i) cur[i] = raw_mask[i] | flooded[i] | tmp[i];
        median_binary(cur.data(), last_fg.data(), W, H, median_k);
        morph_rect(last_fg.data(), dil.data(), W, H, 3, true);
        for(size_t i = 0; i < npx; ++i) { blinks[i] &= dil_inv[i]; dil_inv[i] = (uchar)~dil[i]; blinks[i] &= dil_inv[i]; }
        std::memcpy(fgmask, last_fg.data(), npx);
        {
            const double a1 = (double)(1.0f - aLT), b1 = (1.0 / 255) * (double)aLT, a2 = (double)(1.0f - aST), b2 = (1.0 / 255) * (double)aST;
            for(size_t i = 0; i < npx; ++i) {
                finLT[i] = (float)((double)finLT[i] * a1 + (double)last_fg[i] * b1);
                finST[i] = (float)((double)finST[i] * a2 + (double)last_fg[i] * b2);
            }
        }
        const float ratio = (float)(roi_count - flat_count) / roi_count;
        const size_t off = (size_t)P.lbsp_threshold_offset;
        if(ratio < LBSPDESC_RATIO_MIN && last_nonflat_ratio < LBSPDESC_RATIO_MIN) {
            for(size_t t = 0; t < 256; ++t) if(lut[t] > sat_u8(((float)off + (float)t * P.rel_lbsp_threshold) / 4)) --lut[t];
        } else if(ratio > LBSPDESC_RATIO_MAX && last_nonflat_ratio > LBSPDESC_RATIO_MAX) {
            for(size_t t = 0; t < 256; ++t) if(lut[t] < sat_u8((float)off + 255 * P.rel_lbsp_threshold)) ++lut[t];
        }
        last_nonflat_ratio = ratio;
        // frame-level analysis (:1474-1516)
        ds_area(img, C, ds_frame.data());
        const float bLT = 1.0f - aLT, bST = 1.0f - aST;
        for(size_t i = 0; i < dsLT.size(); ++i) {
            const float sLT = (float)ds_frame[i] * aLT, dLT = dsLT[i] * bLT; dsLT[i] = sLT + dLT;
            const float sST = (float)ds_frame[i] * aST, dST = dsST[i] * bST; dsST[i] = sST + dST;
        }
        const float l1ratio = masked_l1(dsLT.data(), dsST.data(), ds_roi.data(), false) / (float)ds_roi_count;
        if(!auto_reset && l1ratio >= FRAMELEVEL_MIN_L1DIST_THRES * 2) auto_reset = true;
        if(auto_reset || moving_camera) {
            if((frame_idx % BOOTSTRAP_WIN) == 0) {
                std::vector<uchar> bg(npx * C), dsbg((size_t)dsW * dsH * C);
                get_background_image(bg.data());
                ds_area(bg.data(), C, dsbg.data());
                std::vector<float> dsbgf(dsbg.begin(), dsbg.end());
                const float ml1 = masked_l1(dsLT.data(), dsbgf.data(), ds_roi.data(), true) / (float)ds_roi_count;
                const float mcd = C == 1 ? 0.0f : masked_cdist3(dsLT.data(), dsbgf.data(), ds_roi.data()) / (float)ds_roi_count;
                if(moving_camera && ml1 < FRAMELEVEL_MIN_L1DIST_THRES / 4 && mcd < FRAMELEVEL_MIN_CDIST_THRES / 4) {
                    weight_offset = LWORD_WEIGHT_OFFSET; moving_camera = false; refresh_model(1, 1.0f, true);
                } else if(boot && !moving_camera && (ml1 >= FRAMELEVEL_MIN_L1DIST_THRES || mcd >= FRAMELEVEL_MIN_CDIST_THRES)) {
                    weight_offset = 5; moving_camera = true; refresh_model(1, 1.0f, true);
                }
            }
            if(frames_since_reset > BOOTSTRAP_WIN * 2) auto_reset = false;
            else if(l1ratio >= FRAMELEVEL_MIN_L1DIST_THRES && reset_cooldown == 0) {
                frames_since_reset = 0;
                refresh_model(weight_offset / 8, 0.0f, true);
                reset_cooldown = nST;
                std::fill(T.begin(), T.end(), 1.0f);
            } else if(!boot) ++frames_since_reset;
        }
        if(reset_cooldown > 0) --reset_cooldown;
    }

    void apply(const uchar* img, uchar* fgmask, double lr) {
        if(!initialized) throw std::runtime_error("algo & model must be initialized first");
        if(C == 1) apply_impl<1>(img, fgmask, lr); else apply_impl<3>(img, fgmask, lr);
    }

    /// PAWCS.cpp:1525-1557 (weighted mean of the local words' colours, convertTo 8U)
    void get_background_image(uchar* out) const {
        for(size_t p = 0; p < npx; ++p) {
            if(!roi[p]) { for(int c = 0; c < C; ++c) out[p * C + c] = 0; continue; }
            float tw = 0.0f, tc[3] = {0, 0, 0};
            for(int i = 0; i < NW; ++i) {
                const float w = lweight(p * NW + i, frame_idx);
                for(int c = 0; c < C; ++c) tc[c] += (float)lw_color[(p * NW + i) * C + c] * w;
                tw += w;
            }
            for(int c = 0; c < C; ++c) out[p * C + c] = sat_u8(tc[c] / tw);
        }
    }
    /// PAWCS.cpp:1559-1594
    void get_background_desc_image(ushort* out) const {
        for(size_t p = 0; p < npx; ++p) {
            if(!roi[p]) { for(int c = 0; c < C; ++c) out[p * C + c] = 0; continue; }
            float tw = 0.0f, td[3] = {0, 0, 0};
            for(int i = 0; i < NW; ++i) {
                const float w = lweight(p * NW + i, frame_idx);
                for(int c = 0; c < C; ++c) td[c] += (float)lw_desc[(p * NW + i) * C + c] * w;
                tw += w;
            }
            for(int c = 0; c < C; ++c) { const long r = std::lrint((double)(td[c] / tw)); out[p * C + c] = (ushort)(r < 0 ? 0 : (r > 65535 ? 65535 : r)); }
        }
    }

    bool find_buf(const std::string& n, void*& ptr, size_t& bytes) {
#define VB(name, vec) if(n == name) { ptr = (void*)(vec).data(); bytes = (vec).size() * sizeof((vec)[0]); return true; }
        VB("T", T) VB("R", R) VB("v", V) VB("DminLT", DminLT) VB("DminST", DminST) VB("rawLT", rawLT) VB("rawST", rawST) VB("finLT", finLT) VB("finST", finST)
        VB("dsLT", dsLT) VB("dsST", dsST) VB("unstable", unstable) VB("illum", illum) VB("blinks", blinks) VB("lastraw", last_raw) VB("lastrawblink", last_raw_blink)
        VB("dil", dil) VB("dilinv", dil_inv) VB("rawmask", raw_mask)
        VB("lw_first", lw_first) VB("lw_last", lw_last) VB("lw_occ", lw_occ) VB("lw_color", lw_color) VB("lw_desc", lw_desc)
        VB("gw_weight", gw_weight) VB("gw_map", gw_map) VB("gw_bits", gw_bits) VB("gw_color", gw_color) VB("gw_desc", gw_desc) VB("gdict", gdict) VB("glut", glut)
#undef VB
        return false;
    }
    /// scalars 4..9: NW, NG, median_k, weight_offset, moving_camera, last_nonflat_ratio
    void get_scalars(double* d) const { d[4] = NW; d[5] = NG; d[6] = median_k; d[7] = (double)weight_offset; d[8] = moving_camera; d[9] = last_nonflat_ratio; }
    void set_scalars(const double* d) { weight_offset = (size_t)d[7]; moving_camera = d[8] != 0; last_nonflat_ratio = (float)d[9]; }
};

} // namespace lvo
