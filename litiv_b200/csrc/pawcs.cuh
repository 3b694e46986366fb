// litiv_b200 — PAWCS per-frame kernels (replaces BackgroundSubtractorPAWCS_::apply / refreshModel / getBackgroundImage,
// reference video/src/BackgroundSubtractorPAWCS.cpp:86-1603, cited as PAWCS.cpp:line).
//
// Frame = phase A (per pixel: local-dictionary scan + bubble pass, classification, feedback; reads frame-start state of
//          everything the pixel does not own)
//       -> illumination mask of the next frame
//       -> global-dictionary apply (replacement winner, occupancy maps per cell in raster order, fixed-point weights)
//       -> phase B (queued neighbour-dictionary updates, gathered per TARGET pixel in raster order of the source)
//       -> global maintenance (every 8/16 frames) -> bit-packed post-processing (postproc.cuh)
//       -> motion analysis + frame tail (LUT adaptation, reset / moving-camera logic on the device) -> conditional refresh.
// The deterministic parallel semantics ("snapshot" semantics) are specified in DESIGN.md section 2.
//
// HBM layout (per stream; Wp = W rounded up to 32): local words are two sample-major planes [NW][H][Wp]:
// key uint2 (occurrences, first + last mod 2^32) and a record uint4 (colour B,G,R,0 | d0|d1<<16 | d2 | first; 1 channel: uint2
// colour|desc<<16, first): colour, descriptors and `first` always travel together (scan, swaps, new words), and the kernels that
// work one pixel per warp pay one 32-byte sector per word and plane, so one 16-byte record instead of three planes. The weight of a word,
// occ / ((last - first) + 2 (frame - last) + offset) = occ / (2 frame + offset - (first + last)), only needs the key, and every
// frame every pixel re-weights all NW words: 8 B per word in one 64-bit load instead of three 32-bit planes. `first` is touched
// only when a word is matched (last = frame  =>  key.y = first + frame), created or moved; colour / descriptor only while the
// weight sum is below its threshold. State export / import converts to and from (first, last, occurrences). Global words: GDict (small arrays) + occupancy maps [NG][H/2][W/2] f32 +
// per-pixel sort LUT [NG][H][Wp] u8.
#pragma once
#include "subsense.cuh"
#include <cfloat>

namespace lvb {

constexpr int PAW_MAXG = 128;
constexpr uint32_t PAW_BOOTSTRAP = 500u, PAW_WEIGHT_OFFSET = 1000u;
enum { PAW_REQ_NONE = 0, PAW_REQ_REFRESH = 1 };

struct GDict { // global dictionary (indexed by word identity) + the PAWCS frame scalars, device resident
    float weight[PAW_MAXG];
    uint32_t color[PAW_MAXG];          // B,G,R,0 (1ch: byte 0)
    uint2 desc[PAW_MAXG];              // d0|d1<<16, d2 (1ch: d0)
    uint32_t bits[PAW_MAXG];
    int32_t dict[PAW_MAXG];            // dictionary order -> identity (-1: not created yet, init only)
    unsigned long long acc[PAW_MAXG];  // 2^-32 fixed-point weight increments of the current frame
    unsigned long long mapsum[PAW_MAXG];
    uint32_t weight_offset, moving_camera, boot, created;
    float last_nonflat_ratio;
    uint32_t flat_count, rep_winner; int32_t g_rep;
    long long motion_acc, model_l1_acc, model_cd_acc;
    uint32_t refresh_req, refresh_base_occ, refresh_force, set_T_one; float refresh_decr;
    uint32_t ds_roi_count, nST, tail_gate;
    uint32_t wlB_count;                // phase B work-list (reset by the frame tail)
    uint32_t refresh_ticket;           // CTAs of the conditional refresh's last kernel that are done
    uint32_t mflags[PAW_MAXG];         // global maintenance: bit 0 zero the word's map, bit 1 update it
};

struct PawArgs {
    int W, H, Wp, WW, NW, NG, gW, gH;
    size_t plane;
    const uchar* img; size_t ipitch;
    uint2* lw_key; void* lw_rec;   // key = (occurrences, first + last); record = (colour, descriptors, first): PawRec<CH>
    uchar* glut; float* gmap; float* gmap_tmp; GDict* gd;
    float4* maps; float2* fin; void* last_color; void* last_desc;
    const uint32_t* roi_bits; const uint32_t* roi255_bits;
    uint32_t* raw_bits; uint32_t* unstable_bits; const uint32_t* blinks_bits; const uint32_t* lastfg_bits;
    uint32_t* illum_bits; uint32_t* did_bits; const uint32_t* dil_bits; const uint32_t* dilinv_bits;
    uint32_t* intent_bits; uint4* intents; size_t bitplane;
    uint32_t* gop_bits; float* gop_w; uchar* gop_g;
    uint2* hand; uint32_t* wl;   // phase A: scan -> bubble hand-off, work-list of the pixels whose scan goes past PAW_K words
    uint2* wlB;                  // phase B: (target pixel, remaining hits) of the targets whose current hit walks past PAWB_K words
    uchar* lut; FrameCtl* ctl;
    uint64_t seed; uint32_t lr_fixed; int min_color, desc_off;
    int use_tma, collect_stats;
    float rel; int lbsp_off, avg_samples, dsW, dsH;
    const uchar* ds_roi; float* dsLT; float* dsST; uchar* bgimg; // frame-level analysis
};

/// draw % n == 0 for a learning rate n that is 1 for most pixels most of the time (T(x) sits at its floor on a quiet background):
/// the integer division (~20 instructions) is skipped by the warps whose lanes all have n == 1
__device__ __forceinline__ bool paw_draw_hits(uint32_t draw, uint32_t n) { return n == 1u || (draw % n) == 0u; }
/// PAWCS.cpp:1596-1598: occ / ((last - first) + (frame - last) * 2 + offset), all uint32 (wrapping) = occ / (K - (first + last)), K = 2 frame + offset
__device__ __forceinline__ uint32_t paw_wk(uint32_t frame, uint32_t off) { return frame * 2u + off; }
__device__ __forceinline__ float paw_weight(const uint2 key, uint32_t K) { return __fdiv_rn((float)key.x, (float)(K - key.y)); }
__device__ __forceinline__ uint32_t paw_hdist(const uint2& a, const uint2& b) { return __popc(a.x ^ b.x) + __popc((a.y ^ b.y) & 0xFFFFu); }
__device__ __forceinline__ uint32_t paw_hdist(const ushort& a, const ushort& b) { return __popc((uint32_t)(a ^ b)); }
__device__ __forceinline__ uint32_t paw_bits(const uint2& a) { return __popc(a.x) + __popc(a.y & 0xFFFFu); }
__device__ __forceinline__ uint32_t paw_bits(const ushort& a) { return __popc((uint32_t)a); }

/// colour distances of math.hpp: L1dist (u8-wrapping for 3 channels, Q1), cdist :474-496, cmixdist :596-605.
/// Split in two so that callers can stop after the cheap part: cmixdist = (L1 >> 1) + 4 cdist (3 channels) can only be within a
/// threshold if L1 >> 1 is.
template<int CH>
__device__ __forceinline__ uint32_t paw_l1(uint32_t cur, uint32_t bg) {
    if(CH == 1) { const uint32_t a = cur & 0xFFu, b = bg & 0xFFu; return a > b ? a - b : b - a; }
    return __vsadu4(cur & 0x00FFFFFFu, bg & 0x00FFFFFFu) & 0xFFu; // sum of the three byte differences, wrapped like the reference's uchar accumulator
}
/// floor(sqrt(n)) for n < 2^24: == (uint32_t)sqrtf((float)n) of the reference (a correctly rounded root of an integer below 2^24 never
/// rounds up to the next integer), from the approximate root and one correction step
__device__ __forceinline__ uint32_t paw_isqrt(uint32_t n) {
    float s;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"((float)n));
    uint32_t r = (uint32_t)s;
    if(r * r > n) --r; else if((r + 1u) * (r + 1u) <= n) ++r;
    return r;
}
__device__ __forceinline__ uint32_t paw_cdist3(uint32_t cur, uint32_t bg) { // math.hpp:474-496 for uchar triplets
    const uint32_t cu = cur & 0x00FFFFFFu, bu = bg & 0x00FFFFFFu;
    const bool isconst = cu == (cu & 0xFFu) * 0x010101u && bu == (bu & 0xFFu) * 0x010101u; // every channel of both colours equal
    if(isconst || cu == bu) return 0u;
    const uint32_t cs = __dp4a(cu, cu, 0u), bs = __dp4a(bu, bu, 0u), mix = __dp4a(cu, bu, 0u); // sums of squares and the dot product
    // floor(mix^2 / max(bs,1)) exactly, without a 64-bit or double division: q <= cs < 2^18 (Cauchy-Schwarz), so a single-
    // precision estimate (relative error < 2^-21) is off by at most one; the remainder (|r| < 2^19, so 32-bit wrapping arithmetic
    // holds it exactly although mix^2 does not fit) fixes it
    const uint32_t d = max(bs, 1u);
    uint32_t q = (uint32_t)__fdividef(__fmul_rn((float)mix, (float)mix), (float)d);
    const int r = (int)(mix * mix - q * d);
    if(r < 0) --q; else if(r >= (int)d) ++q;
    return paw_isqrt(cs - q);
}
template<int CH>
__device__ __forceinline__ uint32_t paw_color_dist(uint32_t cur, uint32_t bg, uint32_t& l1, uint32_t& cd) {
    l1 = paw_l1<CH>(cur, bg);
    if(CH == 1) { cd = 0; return l1; }
    cd = paw_cdist3(cur, bg);
    return (l1 >> 1) + cd * 4u;
}
/// cmixdist(cur, bg) <= thr, the expensive part only when the cheap part allows it
template<int CH>
__device__ __forceinline__ bool paw_color_within(uint32_t cur, uint32_t bg, uint32_t thr) {
    const uint32_t l1 = paw_l1<CH>(cur, bg);
    if(CH == 1) return l1 <= thr;
    if((l1 >> 1) > thr) return false;
    return (l1 >> 1) + paw_cdist3(cur, bg) * 4u <= thr;
}

/// word record (see the layout note at the top of the file)
template<int CH> struct PawRec;
template<> struct PawRec<3> {
    typedef uint4 T;
    static __device__ __forceinline__ uint32_t col(const T& r) { return r.x; }
    static __device__ __forceinline__ uint2 desc(const T& r) { return make_uint2(r.y, r.z); }
    static __device__ __forceinline__ uint32_t first(const T& r) { return r.w; }
    static __device__ __forceinline__ T make(uint32_t c, uint2 d, uint32_t f) { return make_uint4(c, d.x, d.y, f); }
    static __device__ __forceinline__ void store_cd(T* p, uint32_t c, uint2 d) { *(uint2*)p = make_uint2(c, d.x); ((uint32_t*)p)[2] = d.y; }
    static __device__ __forceinline__ void store_col(T* p, uint32_t c) { ((uint32_t*)p)[0] = c; }
    static __device__ __forceinline__ void store_desc(T* p, uint2 d) { ((uint32_t*)p)[1] = d.x; ((uint32_t*)p)[2] = d.y; }
    static __device__ __forceinline__ uint32_t load_first(const T* p) { return ((const uint32_t*)p)[3]; }
};
template<> struct PawRec<1> {
    typedef uint2 T;
    static __device__ __forceinline__ uchar col(const T& r) { return (uchar)(r.x & 0xFFu); }
    static __device__ __forceinline__ ushort desc(const T& r) { return (ushort)(r.x >> 16); }
    static __device__ __forceinline__ uint32_t first(const T& r) { return r.y; }
    static __device__ __forceinline__ T make(uchar c, ushort d, uint32_t f) { return make_uint2((uint32_t)c | ((uint32_t)d << 16), f); }
    static __device__ __forceinline__ void store_cd(T* p, uchar c, ushort d) { ((uint32_t*)p)[0] = (uint32_t)c | ((uint32_t)d << 16); }
    static __device__ __forceinline__ void store_col(T* p, uchar c) { ((uchar*)p)[0] = c; }
    static __device__ __forceinline__ void store_desc(T* p, ushort d) { ((ushort*)p)[1] = d; }
    static __device__ __forceinline__ uint32_t load_first(const T* p) { return ((const uint32_t*)p)[1]; }
};
#define PAW_REC(A) ((typename PawRec<CH>::T*)(A).lw_rec)

template<int CH> struct PawPlanes {
    typedef typename Pack<CH>::Col Col;
    typedef typename Pack<CH>::Desc Desc;
};
__device__ __forceinline__ uint32_t col_as_u32(const uint32_t& v) { return v; }
__device__ __forceinline__ uint32_t col_as_u32(const uchar& v) { return v; }

/// exchange dictionary positions i and i-1 of one pixel (both planes)
template<int CH>
__device__ __forceinline__ void paw_swap(const PawArgs& A, size_t pix, int i) {
    const size_t a = (size_t)i * A.plane + pix, b = a - A.plane;
    const uint2 k = A.lw_key[a]; const typename PawRec<CH>::T r = PAW_REC(A)[a];
    A.lw_key[a] = A.lw_key[b]; PAW_REC(A)[a] = PAW_REC(A)[b];
    A.lw_key[b] = k; PAW_REC(A)[b] = r;
}

/// search the pixel's sorted global-word LUT (PAWCS.cpp:1073-1079 / :1119-1125); returns the identity or -1.
/// Stage 1 (no dependent loads, all NG LUT bytes in flight): the cheap tests of every position (descriptor bit count, L1 part of the
/// colour distance) give a candidate mask; stage 2 runs the colour-distortion part on the candidates in LUT order.
#ifndef PAW_GFIND
#define PAW_GFIND 4
#endif
template<int CH>
__device__ __forceinline__ int paw_find_gword(const PawArgs& A, const uint32_t* s_gbits, const uint32_t* s_gcolor, size_t pix, uint32_t cur_pack, uint32_t bits, uint32_t thrC, uint32_t thrD) {
    for(int g0 = 0; g0 < A.NG; g0 += PAW_GFIND) {
        uint32_t gv[PAW_GFIND];
#pragma unroll
        for(int k = 0; k < PAW_GFIND; ++k) gv[k] = g0 + k < A.NG ? (uint32_t)A.glut[(size_t)(g0 + k) * A.plane + pix] : 0u;
        uint32_t cand = 0;
#pragma unroll
        for(int k = 0; k < PAW_GFIND; ++k) {
            const uint32_t gb = s_gbits[gv[k]];
            const uint32_t l1 = paw_l1<CH>(cur_pack, s_gcolor[gv[k]]);
            const bool ok = g0 + k < A.NG && (bits > gb ? bits - gb : gb - bits) <= thrD / 4u && (CH == 1 ? l1 : (l1 >> 1)) <= thrC;
            cand |= ok ? (1u << k) : 0u;
        }
        while(cand) {
            const int k = __ffs(cand) - 1;
            cand &= cand - 1u;
            uint32_t g = gv[0];
#pragma unroll
            for(int j = 1; j < PAW_GFIND; ++j) g = k == j ? gv[j] : g;
            if(CH == 1) return (int)g;
            if((paw_l1<CH>(cur_pack, s_gcolor[g]) >> 1) + paw_cdist3(cur_pack, s_gcolor[g]) * 4u <= thrC) return (int)g;
        }
    }
    return -1;
}

// ------------------------------------------------------------------------------------------------------------
// Phase A = scan -> scan tail -> bubble (NW <= PAW_SPLIT_MAX_NW: the hand-off word holds one bit per word).
//
// The reference interleaves, per pixel, the word scan ("while the weight sum is below its threshold") with one bubble-sort pass
// over ALL words (PAWCS.cpp:1002-1065). Step i of that pass looks at position i, which still holds the frame-start word i
// (a swap at step i-1 exchanges positions i-1 and i-2 only), and compares the PRE-update weight of word i with the pre-update
// weight of the word carried from below. So the permutation is a pure function of the frame-start counters, and the scan can
// run without swaps if the words it matched are recorded:
//   pawcs_scan       tile kernel: LBSP + the first PAW_K words (their colour / descriptor / counters are in flight before the
//                    input tile lands). A pixel that ends its scan within them (> 95 %) is classified and gets its feedback step
//                    here; the others (foreground scans all NW words) go to a work-list and NOTHING of theirs is written yet.
//   pawcs_scan_tail  one work-list pixel per lane, the same per-pixel code from word 0 with no depth limit.
//   pawcs_bubble     every pixel, uniform: all NW counters (in flight in chunks), weights, the carry chain of the bubble pass
//                    with the carried word kept in registers (a run of k swaps = k + 1 word moves instead of 2k), the counter
//                    updates of the matched words (hand-off mask), the new word over the last one (:1142-1153).
// Hand-off per pixel (uint2): matched-word mask bits 0..55, flags in the top byte.
// ------------------------------------------------------------------------------------------------------------
#ifndef PAW_KW
#define PAW_KW 4
#endif
constexpr int PAW_K = PAW_KW;
constexpr int PAW_SPLIT_MAX_NW = 56;
constexpr uint32_t PAW_H_OCC = 1u << 24, PAW_H_NEW = 1u << 25, PAW_H_FLAT = 1u << 26, PAW_H_SKIP = 1u << 27; // hand.y: mask bits 32..55 in bits 0..23; SKIP: pawcs_scan_tail owns the pixel

template<int CH> struct PawPix {
    typedef typename Pack<CH>::Col Col;
    typedef typename Pack<CH>::Desc Desc;
    Lookup16 L[CH];
    Col cur_pack; Desc intra_pack;
    uint32_t cur32, bits, thrC, thrD, rate, pixid, frame, woff, wk;   // wk = 2 frame + weight offset (paw_wk)
    float wthr;
    bool unst, flat, border, lastfg, blink, boot, moving;
    size_t pix;
};
struct PawScan { float sum; uint32_t minColor, minDesc, illum_cur, mlo, mhi, scanned; bool did; };
struct PawMaps { float T, R, V, DminLT, DminST, rawLT, rawST, finLT, finST; };
struct PawBits { bool seg, unstable_new, has_intent, has_gop; int intent_row; uint32_t hand_y; };

/// thresholds, update rate and the scan's weight threshold of one pixel (:967-994); key0 = frame-start key of word 0
template<int CH, bool T7>
__device__ __forceinline__ void paw_pix_setup(const PawArgs& A, const uchar* s_lut, PawPix<CH>& P, const uint32_t (&cur)[CH], const PawMaps& M,
                                              const uint2 key0) {
    const FrameCtl* ctl = A.ctl; const GDict* gd = A.gd;
    P.frame = ctl->frame_idx; P.woff = gd->weight_offset; P.boot = gd->boot != 0; P.moving = gd->moving_camera != 0;
    P.wk = paw_wk(P.frame, P.woff);
    uint32_t intra[CH];
#pragma unroll
    for(int c = 0; c < CH; ++c) intra[c] = lbsp_threshold<T7>(P.L[c], cur[c], s_lut[cur[c]]);
    if constexpr (CH == 1) { P.cur_pack = (uchar)cur[0]; P.intra_pack = (ushort)intra[0]; }
    else { P.cur_pack = cur[0] | (cur[1] << 8) | (cur[2] << 16); P.intra_pack = make_uint2(intra[0] | (intra[1] << 16), intra[2]); }
    P.cur32 = col_as_u32(P.cur_pack);
    P.bits = paw_bits(P.intra_pack);
    P.flat = P.bits < (CH == 1 ? 2u : 4u);
    P.rate = A.lr_fixed ? A.lr_fixed : (P.flat ? (uint32_t)ceilf(__fadd_rn(M.T, 1.0f)) / 2u : (uint32_t)ceilf(M.T)); // :987-989
    const uint32_t cbase = (uint32_t)__fmul_rn(__fsqrt_rn(M.R), (float)A.min_color);
    P.thrC = CH == 1 ? cbase / 2u : cbase * 3u;
    const uint32_t dbase = (1u << (uint32_t)floorf(__fadd_rn(M.R, 0.5f))) + (uint32_t)A.desc_off + (P.unst ? (uint32_t)A.desc_off : 0u);
    P.thrD = CH == 1 ? dbase : dbase * 3u;
    P.wthr = __fdiv_rn(paw_weight(key0, P.wk), __fmul_rn(M.R, 2.0f)); // :967-968
}

/// one step of the word scan (:1002-1043) without the swap and without the counter updates (pawcs_bubble applies them)
template<int CH, bool T7, bool DEFER>
__device__ __forceinline__ void paw_test_word(const PawArgs& A, const uchar* s_lut, const PawPix<CH>& P, PawScan& S, int i,
                                              const typename Pack<CH>::Col bc, const typename Pack<CH>::Desc bd, const uint2 key, uint32_t& deferred) {
    typedef typename Pack<CH>::Col Col;
    typedef typename Pack<CH>::Desc Desc;
    ++S.scanned;
    const uint32_t l1 = paw_l1<CH>(P.cur32, col_as_u32(bc));
    if((CH == 1 ? l1 : (l1 >> 1)) > P.thrC) return; // neither the illumination update nor a match is possible
    const uint32_t mix = CH == 1 ? l1 : (l1 >> 1) + paw_cdist3(P.cur32, col_as_u32(bc)) * 4u;
    if(mix > P.thrC) return;
    const uint32_t ihd = paw_hdist(P.intra_pack, bd);
    if((!P.unst || P.flat || P.border) && l1 >= P.thrC / 2u && ihd <= P.thrD / 2u) { // illumination update (:1014-1030)
        const uint32_t mod = S.illum_cur ? (P.rate / 2u + 1u) : P.rate;
        if((philox_draw(A.seed, P.frame, P.pixid, 4u + (uint32_t)i, DOM_PAWCS_A) % mod) == 0u) {
            if(DEFER) deferred |= 1u << i;
            else PawRec<CH>::store_cd(PAW_REC(A) + (size_t)i * A.plane + P.pix, P.cur_pack, P.intra_pack);
            S.did = true; S.illum_cur = 2u;
        }
    }
    uint32_t ehd = 0;
#pragma unroll
    for(int c = 0; c < CH; ++c) {
        const uint32_t b = col_get(bc, c);
        ehd += __popc(lbsp_threshold<T7>(P.L[c], b, s_lut[b]) ^ desc_get(bd, c));
    }
    const uint32_t dd = (ihd + ehd) >> 1;
    if(dd <= P.thrD) {
        S.sum = __fadd_rn(S.sum, paw_weight(key, P.wk));
        if(i < 32) S.mlo |= 1u << i; else S.mhi |= 1u << (i - 32);
        S.minColor = min(S.minColor, mix); S.minDesc = min(S.minDesc, dd);
    }
}

/// classification, global-word look-up, queued neighbour update, feedback (:1070-1269) and the hand-off word of a pixel whose scan ended
template<int CH, bool GPRE = false, bool WRITE_HAND = true>
__device__ __forceinline__ PawBits paw_finish(const PawArgs& A, const uint32_t* s_gbits, const uint32_t* s_gcolor, const PawPix<CH>& P, const PawScan& S,
                                              PawMaps M, int x, int y, int g_pre = -1) {
    const float aLT = A.ctl->aLT, aST = A.ctl->aST;
    const uint32_t colorRange = CH == 1 ? 255u : 765u, descRange = CH == 1 ? 16u : 48u;
    PawBits B; B.seg = false; B.has_intent = false; B.has_gop = false; B.intent_row = 0; B.hand_y = 0;
    bool new_word = false;
    const float sum = S.sum, wthr = P.wthr;
    const uint32_t rate = P.rate;
    const uint4 rnd = philox_block(A.seed, P.frame, P.pixid, 0, DOM_PAWCS_A);
    const float oneLT = __fsub_rn(1.0f, aLT), oneST = __fsub_rn(1.0f, aST);
    const float baseMin = fmaxf(__fdiv_rn((float)S.minColor, (float)colorRange), __fdiv_rn((float)S.minDesc, (float)descRange));
    const size_t cell = (size_t)(y >> 1) * A.gW + (x >> 1);
    if(sum >= wthr || P.border) { // background (:1070-1106)
        M.DminLT = __fadd_rn(__fmul_rn(M.DminLT, oneLT), __fmul_rn(baseMin, aLT));
        M.DminST = __fadd_rn(__fmul_rn(M.DminST, oneST), __fmul_rn(baseMin, aST));
        M.rawLT = __fmul_rn(M.rawLT, oneLT); M.rawST = __fmul_rn(M.rawST, oneST);
        if(paw_draw_hits(rnd.x, rate)) {
            const int g = GPRE ? g_pre : paw_find_gword<CH>(A, s_gbits, s_gcolor, P.pix, P.cur32, P.bits, P.thrC, P.thrD);
            const uint32_t rep = rate == 1u ? (rnd.y & 1u) : rate >= 0x40000000u ? rnd.y : rnd.y % (rate * 2u);
            if(g >= 0 || rep == 0u) {
                A.gop_g[P.pix] = g >= 0 ? (uchar)g : (uchar)0xFE; A.gop_w[P.pix] = sum; B.has_gop = true;
                if(g < 0) atomicMin(&A.gd->rep_winner, P.pixid);
            }
        }
    } else { // foreground (:1107-1155)
        const float nmin = fmaxf(baseMin, __fdiv_rn(__fsub_rn(wthr, sum), wthr));
        M.DminLT = __fadd_rn(__fmul_rn(M.DminLT, oneLT), __fmul_rn(nmin, aLT));
        M.DminST = __fadd_rn(__fmul_rn(M.DminST, oneST), __fmul_rn(nmin, aST));
        M.rawLT = __fadd_rn(__fmul_rn(M.rawLT, oneLT), aLT); M.rawST = __fadd_rn(__fmul_rn(M.rawST, oneST), aST);
        if(P.flat || paw_draw_hits(rnd.x, rate)) {
            const int g = GPRE ? g_pre : paw_find_gword<CH>(A, s_gbits, s_gcolor, P.pix, P.cur32, P.bits, P.thrC, P.thrD);
            if(g < 0) B.seg = true;
            else if(__fadd_rn(sum, __fdiv_rn(A.gmap[(size_t)g * A.gW * A.gH + cell], P.flat ? 2.0f : 4.0f)) < wthr) B.seg = true;
        } else B.seg = true;
        new_word = sum < __fdiv_rn(1.0f, (float)P.woff); // new local word over the last one (:1142-1153): written by pawcs_bubble
    }
    // neighbour dictionary update, queued (:1164-1247)
    if((!B.seg && paw_draw_hits(rnd.z, rate)) || P.border || P.moving) {
        int dx, dy;
        neighbor_offset(!(P.flat || P.border || P.moving), rnd.w, dx, dy);
        const int nx = clampi(x + dx, 2, A.W - 3), ny = clampi(y + dy, 2, A.H - 3);
        if((A.roi_bits[ny * A.WW + (nx >> 5)] >> (nx & 31)) & 1u) {
            A.intents[P.pix] = make_uint4((uint32_t)((ny - y + 2) * 5 + (nx - x + 2)) | (P.thrD << 8), P.thrC, __float_as_uint(wthr), rate);
            B.has_intent = true; B.intent_row = ny - y + 2;
        }
    }
    // feedback (:1252-1269)
    B.unstable_new = (M.R > 3.0f) || (__fsub_rn(M.rawLT, M.finLT) > 0.1f) || (__fsub_rn(M.rawST, M.finST) > 0.1f);
    const float dmin = fminf(M.DminLT, M.DminST), dmax = fmaxf(M.DminLT, M.DminST);
    if(P.lastfg || (dmin < 0.1f && B.seg)) M.T = fminf(__fadd_rn(M.T, __fdiv_rn(0.5f, __fmul_rn(dmax, M.V))), 256.0f);
    else M.T = fmaxf(__fsub_rn(M.T, __fdiv_rn(__fmul_rn(0.25f, M.V), dmax)), 1.0f);
    if(dmax > 0.1f && P.blink) M.V = __fadd_rn(M.V, P.boot ? 2.0f : 1.0f);
    else M.V = fmaxf(__fsub_rn(M.V, __fmul_rn(0.1f, (P.boot || P.flat) ? 2.0f : P.lastfg ? 0.5f : 1.0f)), 0.1f);
    const double rr = (double)__fadd_rn(1.0f, __fmul_rn(dmin, 2.0f));
    if((double)M.R < __dmul_rn(rr, rr)) M.R = __fadd_rn(M.R, __fmul_rn(0.01f, __fsub_rn(M.V, 0.1f)));
    else M.R = fmaxf(__fsub_rn(M.R, __fdiv_rn(0.01f, M.V)), 1.0f);

    A.maps[P.pix * 2] = make_float4(M.T, M.R, M.V, 0.0f);
    A.maps[P.pix * 2 + 1] = make_float4(M.DminLT, M.DminST, M.rawLT, M.rawST);
    ((typename Pack<CH>::Col*)A.last_color)[P.pix] = P.cur_pack;
    ((typename Pack<CH>::Desc*)A.last_desc)[P.pix] = P.intra_pack;
    B.hand_y = S.mhi | ((!P.lastfg || P.moving) ? PAW_H_OCC : 0u) | (new_word ? PAW_H_NEW : 0u) | (P.flat ? PAW_H_FLAT : 0u);
    if(WRITE_HAND) A.hand[P.pix] = make_uint2(S.mlo, B.hand_y);
    return B;
}

#ifndef PAWS_MIN_BLOCKS
#define PAWS_MIN_BLOCKS 4
#endif
template<int CH, bool T7>
__global__ void __launch_bounds__(TILE_W * TILE_H, PAWS_MIN_BLOCKS)
pawcs_scan(const PawArgs A, const __grid_constant__ CUtensorMap tmap) {
    typedef typename Pack<CH>::Col Col;
    typedef typename Pack<CH>::Desc Desc;
    constexpr int PITCH = tile_pitch(CH);
    __shared__ __align__(128) uchar s_tile[PITCH * TILE_ROWS];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ uchar s_lut[256];
    __shared__ uint32_t s_cnt[4];
    __shared__ uint32_t s_gbits[PAW_MAXG], s_gcolor[PAW_MAXG]; // frame-start snapshot of the global words' keys (read-only in this kernel)

    const int x0 = blockIdx.x * TILE_W, y0 = blockIdx.y * TILE_H;
    const int tid = threadIdx.y * TILE_W + threadIdx.x;
    stage_tile_begin<CH>(s_tile, &s_bar, &tmap, A.use_tma, A.img, A.ipitch, A.W, A.H, x0, y0);
    for(int i = tid; i < 256; i += TILE_W * TILE_H) s_lut[i] = A.lut[i];
    for(int i = tid; i < PAW_MAXG; i += TILE_W * TILE_H) { s_gbits[i] = A.gd->bits[i]; s_gcolor[i] = A.gd->color[i]; }
    if(tid < 4) s_cnt[tid] = 0;

    const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
    const bool in_img = (x < A.W) && (y < A.H);
    const bool in_words = y < A.H && (x >> 5) < A.WW;
    const int wi = y * A.WW + (x >> 5);
    const uint32_t lane_bit = 1u << (x & 31);
    uint32_t w_roi = 0, w_roi255 = 0, w_unst = 0, w_blink = 0, w_lastfg = 0, w_illum = 0;
    if(in_words) {
        w_roi = A.roi_bits[wi]; w_roi255 = A.roi255_bits[wi]; w_unst = A.unstable_bits[wi]; w_blink = A.blinks_bits[wi];
        w_lastfg = A.lastfg_bits[wi]; w_illum = A.illum_bits[wi];
    }
    const bool active = in_img && (w_roi & lane_bit);
    const size_t pix = (size_t)y * A.Wp + x;
    float4 m0 = make_float4(0, 0, 0, 0), m1 = m0;
    float2 fin = make_float2(0, 0);
    // the first PAW_K words of the pixel, in flight before the tile lands
    Col bc[PAW_K]; Desc bd[PAW_K]; uint2 wkey[PAW_K];
#pragma unroll
    for(int k = 0; k < PAW_K; ++k) { bc[k] = Col(); bd[k] = Desc(); wkey[k] = make_uint2(0, 0); }
    if(active) {
#pragma unroll
        for(int k = 0; k < PAW_K; ++k)
            if(k < A.NW) {
                const size_t at = (size_t)k * A.plane + pix;
                wkey[k] = A.lw_key[at];
                { const typename PawRec<CH>::T r = PAW_REC(A)[at]; bc[k] = PawRec<CH>::col(r); bd[k] = PawRec<CH>::desc(r); }
            }
        m0 = A.maps[pix * 2]; m1 = A.maps[pix * 2 + 1]; fin = A.fin[pix];
    }
    stage_tile_wait(&s_bar, A.use_tma);

    bool finished = false, did = false, flat = false, unst_old = false;
    PawBits B; B.seg = false; B.unstable_new = false; B.has_intent = false; B.has_gop = false; B.intent_row = 0; B.hand_y = 0;
    uint32_t scanned = 0;
    if(active) {
        PawPix<CH> P;
        P.pix = pix; P.pixid = (uint32_t)(y * A.W + x);
        P.unst = (w_unst & lane_bit) != 0; P.blink = (w_blink & lane_bit) != 0; P.lastfg = (w_lastfg & lane_bit) != 0;
        P.border = !(w_roi255 & lane_bit);
        unst_old = P.unst;
        PawMaps M; M.T = m0.x; M.R = m0.y; M.V = m0.z; M.DminLT = m1.x; M.DminST = m1.y; M.rawLT = m1.z; M.rawST = m1.w; M.finLT = fin.x; M.finST = fin.y;
        uint32_t cur[CH];
        {
            const Window5<CH> Wn = lbsp_window_smem<CH>(s_tile, PITCH, threadIdx.y + HALO, tile_shift(CH) + (int)threadIdx.x * CH);
#pragma unroll
            for(int c = 0; c < CH; ++c) { P.L[c] = lbsp_lookup_window<CH>(Wn, c); cur[c] = win_center<CH>(Wn, c); }
        }
        paw_pix_setup<CH, T7>(A, s_lut, P, cur, M, wkey[0]);
        flat = P.flat;
        PawScan S; S.sum = 0.0f; S.minColor = CH == 1 ? 255u : 765u; S.minDesc = CH == 1 ? 16u : 48u;
        S.illum_cur = (w_illum & lane_bit) ? 1u : 0u; S.mlo = 0; S.mhi = 0; S.scanned = 0; S.did = false;
        uint32_t deferred = 0;
#pragma unroll
        for(int k = 0; k < PAW_K; ++k)
            if(k < A.NW && S.sum < P.wthr) paw_test_word<CH, T7, true>(A, s_lut, P, S, k, bc[k], bd[k], wkey[k], deferred);
        finished = !(PAW_K < A.NW && S.sum < P.wthr);
        if(finished) {
#pragma unroll
            for(int k = 0; k < PAW_K; ++k)
                if((deferred >> k) & 1u) PawRec<CH>::store_cd(PAW_REC(A) + (size_t)k * A.plane + pix, P.cur_pack, P.intra_pack);
            B = paw_finish<CH>(A, s_gbits, s_gcolor, P, S, M, x, y);
            did = S.did; scanned = S.scanned;
        }
    }
    // pixels whose scan goes on: work-list (one atomic per warp)
    {
        const uint32_t um = __ballot_sync(0xFFFFFFFFu, active && !finished);
        if(um) {
            uint32_t base = 0;
            if(threadIdx.x == 0) base = atomicAdd(&A.ctl->wl_count, (uint32_t)__popc(um));
            base = __shfl_sync(0xFFFFFFFFu, base, 0);
            if(active && !finished) { A.wl[base + __popc(um & (lane_bit - 1u))] = (uint32_t)pix; A.hand[pix] = make_uint2(0u, PAW_H_SKIP); }
        }
    }
    const uint32_t b_raw = __ballot_sync(0xFFFFFFFFu, B.seg);
    const uint32_t b_unst = __ballot_sync(0xFFFFFFFFu, finished ? B.unstable_new : unst_old); // an unfinished pixel keeps its old bit for the tail kernel
    const uint32_t b_did = __ballot_sync(0xFFFFFFFFu, did);
    const uint32_t b_gop = __ballot_sync(0xFFFFFFFFu, B.has_gop);
    const uint32_t b_flat = __ballot_sync(0xFFFFFFFFu, flat);
    uint32_t b_int = 0;
#pragma unroll
    for(int d = 0; d < 5; ++d) {
        const uint32_t b = __ballot_sync(0xFFFFFFFFu, B.has_intent && B.intent_row == d);
        if((int)threadIdx.x == d) b_int = b;
    }
    if(in_words) {
        if(threadIdx.x < 5) A.intent_bits[(size_t)threadIdx.x * A.bitplane + wi] = b_int;
        if(threadIdx.x == 0) {
            A.raw_bits[wi] = b_raw; A.unstable_bits[wi] = b_unst; A.did_bits[wi] = b_did; A.gop_bits[wi] = b_gop;
            atomicAdd(&s_cnt[0], __popc(b_flat));
        }
    }
    if(A.collect_stats) {
        uint32_t sc = scanned;
#pragma unroll
        for(int o = 16; o > 0; o >>= 1) sc += __shfl_xor_sync(0xFFFFFFFFu, sc, o);
        if(threadIdx.x == 0) { atomicAdd(&s_cnt[1], sc); atomicAdd(&s_cnt[3], __popc(b_raw)); }
    }
    __syncthreads();
    if(tid == 0) {
        if(s_cnt[0]) atomicAdd(&A.gd->flat_count, s_cnt[0]);
        if(A.collect_stats) {
            atomicAdd(&A.ctl->stat_scanned, (unsigned long long)s_cnt[1]);
            atomicAdd(&A.ctl->stat_fg, (unsigned long long)s_cnt[3]);
        }
    }
}

/// the pixels pawcs_scan could not finish (foreground scans all NW words): one pixel per WARP, nothing of theirs was written yet.
/// Lane j tests word 32 r + j (colour / descriptor distances, weight, the draw of its illumination update); the sequential part of
/// the reference's loop (weight sum in word order until it reaches the threshold, illumination updates whose modulus depends on
/// the earlier ones) is replayed over the ballots of the words that matter; the global-word look-up is one word per lane too.
constexpr int PAW_TAIL_THREADS = 128;
#ifndef PAWT_MIN_BLOCKS
#define PAWT_MIN_BLOCKS 5
#endif
template<int CH, bool T7>
__global__ void __launch_bounds__(PAW_TAIL_THREADS, PAWT_MIN_BLOCKS) pawcs_scan_tail(const PawArgs A) {
    typedef typename Pack<CH>::Col Col;
    typedef typename Pack<CH>::Desc Desc;
    __shared__ uchar s_lut[256];
    __shared__ uint32_t s_gbits[PAW_MAXG], s_gcolor[PAW_MAXG];
    constexpr uint32_t WPB = PAW_TAIL_THREADS / 32;
    const uint32_t n = A.ctl->wl_count;
    if(blockIdx.x * WPB >= n) return;
    for(int i = threadIdx.x; i < 256; i += PAW_TAIL_THREADS) s_lut[i] = A.lut[i];
    for(int i = threadIdx.x; i < PAW_MAXG; i += PAW_TAIL_THREADS) { s_gbits[i] = A.gd->bits[i]; s_gcolor[i] = A.gd->color[i]; }
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31u;
    unsigned long long scanned_acc = 0, fg_acc = 0;
    for(uint32_t e = blockIdx.x * WPB + (threadIdx.x >> 5); e < n; e += gridDim.x * WPB) {
        const size_t pix = A.wl[e];
        const int y = (int)(pix / (size_t)A.Wp), x = (int)(pix - (size_t)y * A.Wp);
        const int wi = y * A.WW + (x >> 5);
        const uint32_t lane_bit = 1u << (x & 31);
        PawPix<CH> P;
        P.pix = pix; P.pixid = (uint32_t)(y * A.W + x);
        P.unst = (A.unstable_bits[wi] & lane_bit) != 0; P.blink = (A.blinks_bits[wi] & lane_bit) != 0; P.lastfg = (A.lastfg_bits[wi] & lane_bit) != 0;
        P.border = !(A.roi255_bits[wi] & lane_bit);
        const float4 m0 = A.maps[pix * 2], m1 = A.maps[pix * 2 + 1]; const float2 fin = A.fin[pix];
        PawMaps M; M.T = m0.x; M.R = m0.y; M.V = m0.z; M.DminLT = m1.x; M.DminST = m1.y; M.rawLT = m1.z; M.rawST = m1.w; M.finLT = fin.x; M.finST = fin.y;
        uint32_t cur[CH];
#pragma unroll
        for(int c = 0; c < CH; ++c) { P.L[c] = lbsp_lookup_smem<CH>(A.img, (int)A.ipitch, x, y, c); cur[c] = A.img[(size_t)y * A.ipitch + (size_t)x * CH + c]; }
        // lane j holds words j and j + 32. All keys at once (the bubble pass below needs every weight); colour / descriptor round by
        // round: a word costs a 32-byte sector per plane here (one pixel per warp), and most pixels of the list stop within the first round
        Col bc[2]; Desc bd[2]; uint2 key[2]; float w[2]; bool have_cd[2]; uint32_t first[2];
#pragma unroll
        for(int t = 0; t < 2; ++t) {
            const int j = (int)lane + 32 * t;
            bc[t] = Col(); bd[t] = Desc(); have_cd[t] = false; first[t] = 0;
            key[t] = j < A.NW ? A.lw_key[(size_t)j * A.plane + pix] : make_uint2(0u, 0u);
        }
        const uint32_t gv_lane = (int)lane < A.NG ? (uint32_t)A.glut[(size_t)lane * A.plane + pix] : 0u; // global-word LUT, position = lane
        paw_pix_setup<CH, T7>(A, s_lut, P, cur, M, make_uint2(__shfl_sync(0xFFFFFFFFu, key[0].x, 0), __shfl_sync(0xFFFFFFFFu, key[0].y, 0)));
        PawScan S; S.sum = 0.0f; S.minColor = CH == 1 ? 255u : 765u; S.minDesc = CH == 1 ? 16u : 48u;
        S.illum_cur = (A.illum_bits[wi] & lane_bit) ? 1u : 0u; S.mlo = 0; S.mhi = 0; S.scanned = 0; S.did = false;
        bool zero_den = false;
#pragma unroll
        for(int t = 0; t < 2; ++t) {
            const int j = (int)lane + 32 * t;
            w[t] = j < A.NW ? paw_weight(key[t], P.wk) : 0.0f;
            zero_den |= j < A.NW && P.wk == key[t].y;
        }
#pragma unroll
        for(int t = 0; t < 2; ++t) {
            const int base = 32 * t;
            if(base < A.NW && S.sum < P.wthr) { // (warp-uniform condition)
                const int i = base + (int)lane;
                bool match = false, cand = false;
                uint32_t mix = 0, dd = 0, drw = 0;
                if(i < A.NW) {
                    const size_t at = (size_t)i * A.plane + pix;
                    { const typename PawRec<CH>::T r = PAW_REC(A)[at]; bc[t] = PawRec<CH>::col(r); bd[t] = PawRec<CH>::desc(r); first[t] = PawRec<CH>::first(r); have_cd[t] = true; }
                    const uint32_t l1 = paw_l1<CH>(P.cur32, col_as_u32(bc[t]));
                    if((CH == 1 ? l1 : (l1 >> 1)) <= P.thrC) {
                        mix = CH == 1 ? l1 : (l1 >> 1) + paw_cdist3(P.cur32, col_as_u32(bc[t])) * 4u;
                        if(mix <= P.thrC) {
                            const uint32_t ihd = paw_hdist(P.intra_pack, bd[t]);
                            uint32_t ehd = 0;
#pragma unroll
                            for(int c = 0; c < CH; ++c) {
                                const uint32_t b = col_get(bc[t], c);
                                ehd += __popc(lbsp_threshold<T7>(P.L[c], b, s_lut[b]) ^ desc_get(bd[t], c));
                            }
                            dd = (ihd + ehd) >> 1;
                            cand = (!P.unst || P.flat || P.border) && l1 >= P.thrC / 2u && ihd <= P.thrD / 2u; // :1014-1030
                            match = dd <= P.thrD;
                        }
                    }
                    if(cand) drw = philox_draw(A.seed, P.frame, P.pixid, 4u + (uint32_t)i, DOM_PAWCS_A);
                }
                const uint32_t mm = __ballot_sync(0xFFFFFFFFu, match), cm = __ballot_sync(0xFFFFFFFFu, cand);
                uint32_t evm = mm | cm, done = 0;
                bool stop = false;
                while(evm && !stop) { // the reference's loop order over the words that do something
                    const int j = __ffs(evm) - 1;
                    evm &= evm - 1u;
                    if((cm >> j) & 1u) {
                        const uint32_t mod = S.illum_cur ? (P.rate / 2u + 1u) : P.rate;
                        if((__shfl_sync(0xFFFFFFFFu, drw, j) % mod) == 0u) { done |= 1u << j; S.did = true; S.illum_cur = 2u; }
                    }
                    if((mm >> j) & 1u) {
                        S.sum = __fadd_rn(S.sum, __shfl_sync(0xFFFFFFFFu, w[t], j));
                        if(t == 0) S.mlo |= 1u << j; else S.mhi |= 1u << j;
                        S.minColor = min(S.minColor, __shfl_sync(0xFFFFFFFFu, mix, j)); S.minDesc = min(S.minDesc, __shfl_sync(0xFFFFFFFFu, dd, j));
                        if(!(S.sum < P.wthr)) { stop = true; S.scanned = (uint32_t)(base + j + 1); }
                    }
                }
                if(!stop) S.scanned = (uint32_t)min(base + 32, A.NW);
                if((done >> lane) & 1u) { // illumination update of word i, in place (the bubble pass below moves it with the rest of the word)
                    const size_t at = (size_t)i * A.plane + pix;
                    PawRec<CH>::store_cd(PAW_REC(A) + at, P.cur_pack, P.intra_pack);
                    bc[t] = P.cur_pack; bd[t] = P.intra_pack;
                }
            }
        }
        // global-word look-up (:1073-1079 / :1119-1125), one LUT position per lane, first hit in LUT order
        int g = -1;
        {
            bool hit = false;
            if((int)lane < A.NG) {
                const uint32_t gb = s_gbits[gv_lane];
                if((P.bits > gb ? P.bits - gb : gb - P.bits) <= P.thrD / 4u) hit = paw_color_within<CH>(P.cur32, s_gcolor[gv_lane], P.thrC);
            }
            const uint32_t hm = __ballot_sync(0xFFFFFFFFu, hit);
            if(hm) g = (int)__shfl_sync(0xFFFFFFFFu, gv_lane, __ffs(hm) - 1);
            for(int gb0 = 32; gb0 < A.NG && g < 0; gb0 += 32) { // more than 32 global words (not with the reference's defaults)
                const int gi = gb0 + (int)lane;
                bool hit2 = false; uint32_t gv = 0;
                if(gi < A.NG) {
                    gv = A.glut[(size_t)gi * A.plane + pix];
                    const uint32_t gb = s_gbits[gv];
                    if((P.bits > gb ? P.bits - gb : gb - P.bits) <= P.thrD / 4u) hit2 = paw_color_within<CH>(P.cur32, s_gcolor[gv], P.thrC);
                }
                const uint32_t hm2 = __ballot_sync(0xFFFFFFFFu, hit2);
                if(hm2) g = (int)__shfl_sync(0xFFFFFFFFu, gv, __ffs(hm2) - 1);
            }
        }
        uint32_t hand_y = 0;
        if(lane == 0) {
            const PawBits B = paw_finish<CH, true, false>(A, s_gbits, s_gcolor, P, S, M, x, y, g);
            hand_y = B.hand_y;
            if(B.seg) atomicOr(&A.raw_bits[wi], lane_bit);
            if(B.unstable_new != P.unst) atomicXor(&A.unstable_bits[wi], lane_bit);
            if(S.did) atomicOr(&A.did_bits[wi], lane_bit);
            if(B.has_gop) atomicOr(&A.gop_bits[wi], lane_bit);
            if(B.has_intent) atomicOr(&A.intent_bits[(size_t)B.intent_row * A.bitplane + wi], lane_bit);
            scanned_acc += S.scanned; fg_acc += B.seg ? 1u : 0u;
        }
        // the pixel's bubble pass (what pawcs_bubble does for the other pixels, which runs beside this kernel). The carry chain
        // "if(w_i > carried) swap else carried = w_i" keeps carried = min(w_0..w_i), so bit i of the swap mask is
        // w_i > min(w_0..w_{i-1}): an exclusive prefix minimum over the warp's two slots (no NaN: every denominator is non-zero,
        // checked; otherwise the chain is replayed step by step). Every word that moves or was matched is stored at its final position
        // once all lanes hold theirs in registers.
        {
            hand_y = __shfl_sync(0xFFFFFFFFu, hand_y, 0);
            const unsigned long long matched = ((unsigned long long)(hand_y & 0x00FFFFFFu) << 32) | S.mlo;
            const bool occ_en = (hand_y & PAW_H_OCC) != 0;
            const uint32_t occ_incr = (1u + A.ctl->cooldown) << ((P.flat || P.boot) ? 1 : 0);
            unsigned long long swaps = 0ull;
            if(__any_sync(0xFFFFFFFFu, zero_den)) {
                float last_w = FLT_MAX;
                for(int i = 0; i < A.NW; ++i) { // :1044-1052 (warp-uniform)
                    const float wi_ = __shfl_sync(0xFFFFFFFFu, i < 32 ? w[0] : w[1], i & 31);
                    if(wi_ > last_w) swaps |= 1ull << i; else last_w = wi_;
                }
            } else {
                float sA = w[0], sB = (int)lane + 32 < A.NW ? w[1] : FLT_MAX;
#pragma unroll
                for(int d = 1; d < 32; d <<= 1) {
                    const float ta = __shfl_up_sync(0xFFFFFFFFu, sA, d), tb = __shfl_up_sync(0xFFFFFFFFu, sB, d);
                    if((int)lane >= d) { sA = fminf(sA, ta); sB = fminf(sB, tb); }
                }
                const float totA = __shfl_sync(0xFFFFFFFFu, sA, 31);
                float exA = __shfl_up_sync(0xFFFFFFFFu, sA, 1), exB = __shfl_up_sync(0xFFFFFFFFu, sB, 1);
                if(lane == 0) { exA = FLT_MAX; exB = totA; } else exB = fminf(exB, totA);
                const uint32_t lo = __ballot_sync(0xFFFFFFFFu, (int)lane < A.NW && w[0] > exA);
                const uint32_t hi = __ballot_sync(0xFFFFFFFFu, (int)lane + 32 < A.NW && w[1] > exB);
                swaps = ((unsigned long long)hi << 32) | lo;
            }
            bool moved[2], upd[2]; size_t dst[2];
#pragma unroll
            for(int t = 0; t < 2; ++t) {
                const int j = (int)lane + 32 * t;
                moved[t] = false; upd[t] = false; dst[t] = 0;
                if(j < A.NW) {
                    const size_t at = (size_t)j * A.plane + pix;
                    const bool down = (swaps >> j) & 1ull;
                    const int pos = down ? j - 1 : j + (__ffsll((long long)~(swaps >> (j + 1))) - 1); // carried up through the run of swaps that follows
                    moved[t] = pos != j; upd[t] = (matched >> j) & 1ull;
                    dst[t] = (size_t)pos * A.plane + pix;
                    if(moved[t] && !have_cd[t]) { const typename PawRec<CH>::T r = PAW_REC(A)[at]; bc[t] = PawRec<CH>::col(r); bd[t] = PawRec<CH>::desc(r); first[t] = PawRec<CH>::first(r); }
                    else if(upd[t] && !have_cd[t]) first[t] = PawRec<CH>::load_first(PAW_REC(A) + at);
                    if(upd[t]) key[t] = make_uint2((occ_en && w[t] < 1.0f) ? key[t].x + occ_incr : key[t].x, first[t] + P.frame); // :1035-1038
                }
            }
            __syncwarp();
#pragma unroll
            for(int t = 0; t < 2; ++t) {
                if(moved[t]) { A.lw_key[dst[t]] = key[t]; PAW_REC(A)[dst[t]] = PawRec<CH>::make(bc[t], bd[t], first[t]); }
                else if(upd[t]) A.lw_key[dst[t]] = key[t];
            }
            __syncwarp();
            if((hand_y & PAW_H_NEW) && lane == 0) { // new local word over the last one (:1142-1153)
                const size_t at = (size_t)(A.NW - 1) * A.plane + pix;
                PAW_REC(A)[at] = PawRec<CH>::make(P.cur_pack, P.intra_pack, P.frame);
                A.lw_key[at] = make_uint2(occ_incr, P.frame * 2u);
            }
        }
    }
    if(A.collect_stats && lane == 0) {
        if(scanned_acc) atomicAdd(&A.ctl->stat_scanned, scanned_acc);
        if(fg_acc) atomicAdd(&A.ctl->stat_fg, fg_acc);
    }
}

/// bubble pass over all words of every pixel + counter updates of the matched words + the new word (see the block comment above).
/// Two steps per pixel: (1) uniform, no stores: all NW keys (PAWU_CHUNK 64-bit loads in flight) and the carry chain
/// "if(w_i > carried) swap" of the reference, which yields the swap mask s (bit i: word i moves down to i-1);
/// (2) sparse: only the words that move (s | s >> 1: a run of k swaps rotates k + 1 words, the carried one in registers) or were
/// matched by the scan (hand-off mask: last = frame, occurrences += incr) are read again (L1 / L2 hits) and stored.
///
/// Step (1) compares the FLOAT quotients fl(o_i / d_i) > fl(o_c / d_c) without dividing: for operands below 2^24 (exact in float)
/// rounding is monotone, so o_i d_c <= o_c d_i (64-bit products) means "no swap", and a relative gap above 2^-22 means the rounded
/// quotients differ too ("swap"); the sliver in between (near ties) takes the two IEEE divisions. A pixel with a counter >= 2^24 or a
/// zero denominator anywhere redoes step (1) with the divisions (paw_swaps_by_division).
#ifndef PAWU_CHUNK
#define PAWU_CHUNK 10
#endif
#ifndef PAWU_MIN_BLOCKS
#define PAWU_MIN_BLOCKS 5
#endif
/// near tie of two word weights: the reference's comparison of the rounded quotients (kept out of line: it is rare, and inlined the two
/// IEEE divisions are if-converted into every step of the chain)
__device__ __noinline__ bool paw_near_tie(uint32_t oi, uint32_t di, uint32_t oc, uint32_t dc) {
    return __fdiv_rn((float)oi, (float)di) > __fdiv_rn((float)oc, (float)dc);
}
__device__ __noinline__ unsigned long long paw_swaps_by_division(const uint2* pk, size_t plane, int NW, uint32_t K) {
    unsigned long long swaps = 0ull;
    float last_w = FLT_MAX;
    for(int i = 0; i < NW; ++i) {
        const float w = paw_weight(pk[(size_t)i * plane], K);
        if(w > last_w) swaps |= 1ull << i; else last_w = w; // :1044-1052
    }
    return swaps;
}
template<int CH>
__global__ void __launch_bounds__(256, PAWU_MIN_BLOCKS) pawcs_bubble(const PawArgs A) {
    typedef typename Pack<CH>::Col Col;
    typedef typename Pack<CH>::Desc Desc;
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if(x >= A.W || y >= A.H) return;
    if(!((A.roi_bits[y * A.WW + (x >> 5)] >> (x & 31)) & 1u)) return;
    const size_t pix = (size_t)y * A.Wp + x;
    const uint2 hand = A.hand[pix];
    if(hand.y & PAW_H_SKIP) return; // pawcs_scan_tail does the bubble pass of its pixels itself
    const FrameCtl* ctl = A.ctl; const GDict* gd = A.gd;
    const uint32_t frame = ctl->frame_idx, K = paw_wk(frame, gd->weight_offset);
    constexpr int CHUNK = PAWU_CHUNK;
    unsigned long long swaps = 0ull;
    {
        uint32_t oc = 0xFFFFFFFFu, dc = 1u, unsafe = 0u; // carried word; the initial value never lets word 0 swap
        const uint2* pk = A.lw_key + pix;
        int i0 = 0;
        for(; i0 + CHUNK <= A.NW; i0 += CHUNK) {
            uint2 key[CHUNK];
#pragma unroll
            for(int k = 0; k < CHUNK; ++k) key[k] = pk[(size_t)k * A.plane];
            uint32_t sw = 0;
#pragma unroll
            for(int k = 0; k < CHUNK; ++k) {
                const uint32_t oi = key[k].x, di = K - key[k].y;
                unsafe |= oi | (di - 1u);
                const unsigned long long a = (unsigned long long)oi * dc, b = (unsigned long long)oc * di;
                bool s = a > b;
                if(s && !((a - b) > (b >> 22))) s = paw_near_tie(oi, di, oc, dc);
                sw |= s ? (1u << k) : 0u;
                oc = s ? oc : oi; dc = s ? dc : di;
            }
            swaps |= (unsigned long long)sw << i0;
            pk += (size_t)CHUNK * A.plane;
        }
        for(; i0 < A.NW; ++i0) { // NW not a multiple of the chunk
            const uint2 key = *pk;
            const uint32_t oi = key.x, di = K - key.y;
            unsafe |= oi | (di - 1u);
            const unsigned long long a = (unsigned long long)oi * dc, b = (unsigned long long)oc * di;
            bool s = a > b;
            if(s && !((a - b) > (b >> 22))) s = paw_near_tie(oi, di, oc, dc);
            swaps |= s ? (1ull << i0) : 0ull;
            oc = s ? oc : oi; dc = s ? dc : di;
            pk += A.plane;
        }
        if(unsafe >> 24) swaps = paw_swaps_by_division(A.lw_key + pix, A.plane, A.NW, K);
    }
    const bool occ_en = (hand.y & PAW_H_OCC) != 0;
    const uint32_t occ_incr = (1u + ctl->cooldown) << (((hand.y & PAW_H_FLAT) || gd->boot) ? 1 : 0);
    const unsigned long long matched = ((unsigned long long)(hand.y & 0x00FFFFFFu) << 32) | hand.x;
    unsigned long long ev = matched | swaps | (swaps >> 1);
    uint2 ckey = make_uint2(0, 0); typename PawRec<CH>::T crec = typename PawRec<CH>::T(); // the carried word of the current run
    // events in word order; the next event's word is fetched while the current one is stored (an event at i writes positions i-1
    // and i only, the next one reads a position > i)
    struct Ev { int i; uint2 key; typename PawRec<CH>::T rec; };
    auto fetch = [&](Ev& e) {
        e.i = __ffsll((long long)ev) - 1;
        if(e.i < 0) return;
        ev &= ev - 1ull;
        const size_t at = (size_t)e.i * A.plane + pix;
        e.key = A.lw_key[at]; e.rec = PAW_REC(A)[at];
    };
    Ev cur, nxt;
    fetch(cur);
    while(cur.i >= 0) {
        fetch(nxt);
        const int i = cur.i;
        const size_t at = (size_t)i * A.plane + pix;
        uint2 key = cur.key;
        const bool down = (swaps >> i) & 1ull, next_down = (swaps >> (i + 1)) & 1ull;
        if((matched >> i) & 1ull) { // :1035-1038: last = frame, occurrences += incr while the (old) weight is below 1
            const float w = paw_weight(key, K);
            key = make_uint2((occ_en && w < 1.0f) ? key.x + occ_incr : key.x, PawRec<CH>::first(cur.rec) + frame);
        }
        if(down) { // word i -> position i-1; the run ends here if word i+1 stays
            const size_t ab = at - A.plane;
            PAW_REC(A)[ab] = cur.rec; A.lw_key[ab] = key;
            if(!next_down) { A.lw_key[at] = ckey; PAW_REC(A)[at] = crec; }
        } else if(next_down) { // word i is carried up through the run that starts at i+1
            ckey = key; crec = cur.rec;
        } else A.lw_key[at] = key; // matched, stays where it is
        cur = nxt;
    }
    if(hand.y & PAW_H_NEW) { // new local word over the last one (:1142-1153)
        const size_t at = (size_t)(A.NW - 1) * A.plane + pix;
        PAW_REC(A)[at] = PawRec<CH>::make(((const Col*)A.last_color)[pix], ((const Desc*)A.last_desc)[pix], frame);
        A.lw_key[at] = make_uint2(occ_incr, frame * 2u);
    }
}

/// global dictionary, step 1: the first pixel (raster order) that asked for it replaces the last word of the dictionary
/// (gridDim.x CTAs zero the word's occupancy map, thread 0 of CTA 0 rewrites the word; nobody reads what it writes in this kernel)
template<int CH>
__global__ void __launch_bounds__(1024) pawcs_gword_replace(const PawArgs A) {
    GDict* gd = A.gd;
    const uint32_t win = gd->rep_winner;
    if(win == 0xFFFFFFFFu) return;
    const int g = gd->dict[A.NG - 1];
    float* m = A.gmap + (size_t)g * A.gW * A.gH;
    for(int i = blockIdx.x * blockDim.x + threadIdx.x; i < A.gW * A.gH; i += gridDim.x * blockDim.x) m[i] = 0.0f;
    if(threadIdx.x == 0 && blockIdx.x == 0) {
        const int x = (int)(win % (uint32_t)A.W), y = (int)(win / (uint32_t)A.W);
        const size_t pix = (size_t)y * A.Wp + x;
        if constexpr (CH == 1) {
            const ushort d = ((const ushort*)A.last_desc)[pix];
            gd->color[g] = ((const uchar*)A.last_color)[pix]; gd->desc[g] = make_uint2(d, 0); gd->bits[g] = paw_bits(d);
        } else {
            const uint2 d = ((const uint2*)A.last_desc)[pix];
            gd->color[g] = ((const uint32_t*)A.last_color)[pix]; gd->desc[g] = d; gd->bits[g] = paw_bits(d);
        }
        gd->weight[g] = 0.0f; gd->g_rep = g;
    }
}
/// step 2: occupancy updates, one thread per map cell, its <=4 pixels in raster order (PAWCS.cpp:1098-1102)
__global__ void __launch_bounds__(256) pawcs_gword_apply(const PawArgs A) {
    __shared__ unsigned long long s_acc[PAW_MAXG];
    GDict* gd = A.gd;
    for(int i = threadIdx.x; i < A.NG; i += blockDim.x) s_acc[i] = 0ull;
    __syncthreads();
    const int cx = blockIdx.x * 32 + (threadIdx.x & 31), cy = blockIdx.y * 8 + (threadIdx.x >> 5);
    if(cx < A.gW && cy < A.gH) {
        const uint32_t win = gd->rep_winner;
        const size_t cell = (size_t)cy * A.gW + cx, msz = (size_t)A.gW * A.gH;
#pragma unroll
        for(int k = 0; k < 4; ++k) {
            const int x = cx * 2 + (k & 1), y = cy * 2 + (k >> 1);
            if(x >= A.W || y >= A.H) continue;
            if(!((A.gop_bits[y * A.WW + (x >> 5)] >> (x & 31)) & 1u)) continue;
            const size_t pix = (size_t)y * A.Wp + x;
            int g = A.gop_g[pix];
            if(g == 0xFE) { if((uint32_t)(y * A.W + x) != win) continue; g = gd->g_rep; }
            const float w = A.gop_w[pix];
            float* m = A.gmap + (size_t)g * msz + cell;
            const float cw = *m;
            if(cw < w) { *m = __fadd_rn(cw, w); atomicAdd(&s_acc[g], (unsigned long long)__double2ll_rn((double)w * 4294967296.0)); }
        }
    }
    __syncthreads();
    for(int i = threadIdx.x; i < A.NG; i += blockDim.x) if(s_acc[i]) atomicAdd(&gd->acc[i], s_acc[i]);
}
/// step 3 (1 CTA): fold the fixed-point increments into the float weights
__device__ __forceinline__ void paw_gdict_bubble_pass(const PawArgs& A) { // :1316-1317
    GDict* gd = A.gd;
    for(int i = 1; i < A.NG; ++i)
        if(gd->weight[gd->dict[i]] > gd->weight[gd->dict[i - 1]]) { const int t = gd->dict[i]; gd->dict[i] = gd->dict[i - 1]; gd->dict[i - 1] = t; }
}
__global__ void __launch_bounds__(128) pawcs_gword_finish(const PawArgs A, int then_bubble) {
    GDict* gd = A.gd;
    const int g = threadIdx.x;
    if(g < A.NG) {
        const unsigned long long a = gd->acc[g];
        if(a) { gd->weight[g] = (float)((double)gd->weight[g] + (double)(long long)a / 4294967296.0); gd->acc[g] = 0ull; }
    }
    if(g == 0) gd->rep_winner = 0xFFFFFFFFu;
    if(then_bubble) { // no maintenance this frame: the dictionary bubble pass (:1316-1317) follows at once
        __syncthreads();
        if(g == 0) paw_gdict_bubble_pass(A);
    }
}

/// Phase B: queued neighbour-dictionary updates (PAWCS.cpp:1164-1247), gathered per TARGET pixel, raster order of the source.
/// pawcs_phaseB: one target per thread. The sources of the tile + 2-px halo mark their targets in shared memory (bit = position of
/// the source in the target's 5x5 window, ascending = raster order), every target pops its hits in order. A hit walks the target's
/// dictionary while its weight sum is below the source's threshold; here only the first PAWB_K words (in flight together), with
/// the word updates held back until the walk is known to end within them. A target whose current hit needs more words is pushed to
/// a list with its remaining hits and finished by pawcs_phaseB_tail: one target per WARP, one word per lane, the sequential part
/// (weight sum in word order) replayed over the ballot of the credited words.
#ifndef PAWB_MIN_BLOCKS
#define PAWB_MIN_BLOCKS 4
#endif
#ifndef PAWB_KW
#define PAWB_KW 4
#endif
constexpr int PAWB_K = PAWB_KW;
template<int CH> struct PawHit {
    typename Pack<CH>::Col sc; typename Pack<CH>::Desc sd, td;
    uint32_t sc32, thrC, thrD, rate, src_id, occ_incr;
    float wthr;
    bool sflat, tflat, traw;
};
/// the source pixel (qx,qy) updates the dictionary of target (x,y) with its own colour / descriptor / thresholds
template<int CH>
__device__ __forceinline__ void paw_hit_load(const PawArgs& A, PawHit<CH>& Hh, int x, int y, int qx, int qy, uint32_t cooldown, bool boot) {
    typedef typename Pack<CH>::Desc Desc;
    const uint32_t flatK = CH == 1 ? 2u : 4u;
    const size_t qpix = (size_t)qy * A.Wp + qx, pix = (size_t)y * A.Wp + x;
    const uint4 rec = A.intents[qpix];
    Hh.thrD = rec.x >> 8; Hh.thrC = rec.y; Hh.rate = rec.w; Hh.wthr = __uint_as_float(rec.z);
    const uchar* src = A.img + (size_t)qy * A.ipitch + (size_t)qx * CH;
    if constexpr (CH == 1) { Hh.sc = src[0]; Hh.sc32 = src[0]; } else { Hh.sc = (uint32_t)src[0] | ((uint32_t)src[1] << 8) | ((uint32_t)src[2] << 16); Hh.sc32 = Hh.sc; }
    Hh.sd = ((const Desc*)A.last_desc)[qpix];   // == the source's intra descriptor of this frame
    Hh.td = ((const Desc*)A.last_desc)[pix];    // target's intra descriptor of this frame
    Hh.sflat = paw_bits(Hh.sd) < flatK; Hh.tflat = paw_bits(Hh.td) < flatK;
    Hh.occ_incr = (1u + cooldown) << ((Hh.sflat || boot) ? 1 : 0);
    Hh.traw = (A.raw_bits[y * A.WW + (x >> 5)] >> (x & 31)) & 1u;
    Hh.src_id = (uint32_t)(qy * A.W + qx);
}
/// one word of a hit's walk (:1180-1230): bit 0 credit, bit 1 take the source's descriptor, bit 2 take the source's colour
template<int CH>
__device__ __forceinline__ uint32_t paw_hit_word(const PawArgs& A, const PawHit<CH>& Hh, int j, typename Pack<CH>::Col bc, typename Pack<CH>::Desc bd, uint32_t frame, bool boot) {
    uint32_t l1, cd;
    const uint32_t mix = paw_color_dist<CH>(Hh.sc32, col_as_u32(bc), l1, cd);
    const uint32_t hd = paw_hdist(Hh.sd, bd);
    if(mix <= Hh.thrC && hd <= Hh.thrD) return 1u;
    if(!Hh.traw && Hh.sflat && (boot || (philox_draw(A.seed, frame, Hh.src_id, (uint32_t)j, DOM_PAWCS_B) % Hh.rate) == 0u)) {
        const uint32_t lhd = paw_hdist(Hh.sd, Hh.td);
        if(mix <= Hh.thrC && lhd <= Hh.thrD / 2u) return 3u;
        if(CH != 1 && Hh.tflat && lhd + hd <= Hh.thrD && cd <= Hh.thrC / 4u) return 5u;
    }
    return 0u;
}
/// counter / colour / descriptor updates of a credited word (Q8: the 1-channel path of the reference updates a by-value copy, PAWCS.cpp:838)
template<int CH>
__device__ __forceinline__ void paw_hit_commit(const PawArgs& A, const PawHit<CH>& Hh, size_t at, uint32_t flags, uint2 key, uint32_t first, float w, typename Pack<CH>::Desc bd, uint32_t frame) {
    typedef typename Pack<CH>::Col Col;
    typedef typename Pack<CH>::Desc Desc;
    if(CH == 1) return;
    const uint32_t incr = paw_bits(bd) < (CH == 1 ? 2u : 4u) ? Hh.occ_incr * 2u : Hh.occ_incr;
    A.lw_key[at] = make_uint2(w < 1.0f ? key.x + incr : key.x, first + frame); // last = frame
    if(flags & 2u) PawRec<CH>::store_desc(PAW_REC(A) + at, Hh.sd);
    if(flags & 4u) PawRec<CH>::store_col(PAW_REC(A) + at, Hh.sc);
}
template<int CH>
__device__ __forceinline__ void paw_hit_new_word(const PawArgs& A, const PawHit<CH>& Hh, size_t pix, uint32_t frame) {
    const size_t at = (size_t)(A.NW - 1) * A.plane + pix;
    PAW_REC(A)[at] = PawRec<CH>::make(Hh.sc, Hh.sd, frame);
    A.lw_key[at] = make_uint2(Hh.occ_incr, frame * 2u);
}

template<int CH>
__global__ void __launch_bounds__(256, PAWB_MIN_BLOCKS) pawcs_phaseB(const PawArgs A) {
    typedef typename Pack<CH>::Col Col;
    typedef typename Pack<CH>::Desc Desc;
    __shared__ uint32_t s_hits[8][32];
    const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 8;
    const int tid = threadIdx.y * 32 + threadIdx.x;
    s_hits[threadIdx.y][threadIdx.x] = 0u;
    __syncthreads();
    // pass 1: every source of the tile + halo marks its target (bit = window position of the source, raster order)
    for(int sidx = tid; sidx < 36 * 12; sidx += 256) {
        const int sy = y0 - 2 + sidx / 36, sx = x0 - 2 + sidx % 36;
        if(sx < 0 || sy < 0 || sx >= A.W || sy >= A.H) continue;
        const uint32_t code = A.intents[(size_t)sy * A.Wp + sx].x & 0xFFu;
        if(code >= 25u) continue;
        const int row = (int)code / 5, col = (int)code - row * 5;
        if(!((A.intent_bits[(size_t)row * A.bitplane + (size_t)sy * A.WW + (sx >> 5)] >> (sx & 31)) & 1u)) continue; // no intent this frame
        const int tx = sx + col - 2 - x0, ty = sy + row - 2 - y0;
        if(tx < 0 || ty < 0 || tx >= 32 || ty >= 8) continue;
        atomicOr(&s_hits[ty][tx], 1u << ((4 - row) * 5 + (4 - col)));
    }
    __syncthreads();
    const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
    if(x < 2 || y < 2 || x > A.W - 3 || y > A.H - 3) return;
    uint32_t hits = s_hits[threadIdx.y][threadIdx.x];
    if(!hits) return;
    const FrameCtl* ctl = A.ctl; const GDict* gd = A.gd;
    const uint32_t frame = ctl->frame_idx, cooldown = ctl->cooldown, woff = gd->weight_offset, wk = paw_wk(frame, woff);
    const bool boot = gd->boot != 0;
    const size_t pix = (size_t)y * A.Wp + x;
    const float init_w = __fdiv_rn(1.0f, (float)woff);
    // pass 2: every lane pops ITS next hit
    while(hits) {
        const int hi_ = __ffs(hits) - 1;
        const int r_ = hi_ / 5, k = hi_ - r_ * 5;
        PawHit<CH> Hh;
        paw_hit_load<CH>(A, Hh, x, y, x - 2 + k, y + r_ - 2, cooldown, boot);
        Col bc[PAWB_K]; Desc bd[PAWB_K]; uint2 key[PAWB_K]; uint32_t first[PAWB_K];
#pragma unroll
        for(int j = 0; j < PAWB_K; ++j) {
            if(j < A.NW) { const size_t at = (size_t)j * A.plane + pix; const typename PawRec<CH>::T r = PAW_REC(A)[at]; bc[j] = PawRec<CH>::col(r); bd[j] = PawRec<CH>::desc(r); first[j] = PawRec<CH>::first(r); key[j] = A.lw_key[at]; }
            else { bc[j] = Col(); bd[j] = Desc(); key[j] = make_uint2(0, 0); first[j] = 0; }
        }
        float sum = 0.0f, w[PAWB_K]; uint32_t fl[PAWB_K];
#pragma unroll
        for(int j = 0; j < PAWB_K; ++j) {
            fl[j] = 0u; w[j] = 0.0f;
            if(j < A.NW && sum < Hh.wthr) {
                fl[j] = paw_hit_word<CH>(A, Hh, j, bc[j], bd[j], frame, boot);
                if(fl[j] & 1u) { w[j] = paw_weight(key[j], wk); sum = __fadd_rn(sum, w[j]); }
            }
        }
        if(PAWB_K < A.NW && sum < Hh.wthr) { // the walk goes on: this hit and the later ones of this target go to the tail kernel
            const uint32_t e = atomicAdd(&A.gd->wlB_count, 1u);
            A.wlB[e] = make_uint2((uint32_t)pix, hits);
            break;
        }
#pragma unroll
        for(int j = 0; j < PAWB_K; ++j)
            if(fl[j] & 1u) paw_hit_commit<CH>(A, Hh, (size_t)j * A.plane + pix, fl[j], key[j], first[j], w[j], bd[j], frame);
        if(sum < init_w) paw_hit_new_word<CH>(A, Hh, pix, frame);
        hits &= hits - 1u;
    }
}

#ifndef PAWBT_MIN_BLOCKS
#define PAWBT_MIN_BLOCKS 6
#endif
template<int CH>
__global__ void __launch_bounds__(PAW_TAIL_THREADS, PAWBT_MIN_BLOCKS) pawcs_phaseB_tail(const PawArgs A) {
    typedef typename Pack<CH>::Col Col;
    typedef typename Pack<CH>::Desc Desc;
    constexpr uint32_t WPB = PAW_TAIL_THREADS / 32;
    const uint32_t n = A.gd->wlB_count;
    const uint32_t lane = threadIdx.x & 31u;
    const FrameCtl* ctl = A.ctl; const GDict* gd = A.gd;
    const uint32_t frame = ctl->frame_idx, cooldown = ctl->cooldown, woff = gd->weight_offset, wk = paw_wk(frame, woff);
    const bool boot = gd->boot != 0;
    const float init_w = __fdiv_rn(1.0f, (float)woff);
    for(uint32_t e = blockIdx.x * WPB + (threadIdx.x >> 5); e < n; e += gridDim.x * WPB) {
        const uint2 ent = A.wlB[e];
        const size_t pix = ent.x;
        const int y = (int)(pix / (size_t)A.Wp), x = (int)(pix - (size_t)y * A.Wp);
        uint32_t hits = ent.y;
        while(hits) { // (warp-uniform)
            const int hi_ = __ffs(hits) - 1;
            hits &= hits - 1u;
            const int r_ = hi_ / 5, k = hi_ - r_ * 5;
            PawHit<CH> Hh;
            paw_hit_load<CH>(A, Hh, x, y, x - 2 + k, y + r_ - 2, cooldown, boot);
            // a word costs a 32-byte sector per plane here (one target per warp): colour / descriptor round by round, key and first only
            // for the credited words
            float sum = 0.0f;
            bool stop = false;
            for(int base = 0; base < A.NW && !stop; base += 32) { // (warp-uniform)
                const int j = base + (int)lane;
                const size_t at = (size_t)j * A.plane + pix;
                uint32_t fl = 0u, first = 0u; float w = 0.0f; uint2 key = make_uint2(0u, 0u); Desc bd = Desc();
                if(j < A.NW) {
                    const typename PawRec<CH>::T r = PAW_REC(A)[at];
                    bd = PawRec<CH>::desc(r); first = PawRec<CH>::first(r);
                    fl = paw_hit_word<CH>(A, Hh, j, PawRec<CH>::col(r), bd, frame, boot);
                    if(fl & 1u) { key = A.lw_key[at]; w = paw_weight(key, wk); }
                }
                uint32_t m = __ballot_sync(0xFFFFFFFFu, fl & 1u), upto = 0u;
                while(m && !stop) { // the credited words in word order, until the weight sum reaches the source's threshold
                    const int jj = __ffs(m) - 1;
                    m &= m - 1u;
                    sum = __fadd_rn(sum, __shfl_sync(0xFFFFFFFFu, w, jj));
                    upto |= 1u << jj;
                    if(!(sum < Hh.wthr)) stop = true;
                }
                if((upto >> lane) & 1u) paw_hit_commit<CH>(A, Hh, at, fl, key, first, w, bd, frame);
            }
            if(sum < init_w && lane == 0) paw_hit_new_word<CH>(A, Hh, pix, frame);
            __syncwarp();
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// global maintenance (PAWCS.cpp:1300-1334): sum / weights / decay+blur / copy kernels over (map chunks x words), then the dictionary bubble pass, then the per-pixel LUTs
// ------------------------------------------------------------------------------------------------------------
// (one CTA per word took 0.87 ms at 1080p; now every step is spread over (chunks x words) CTAs)
constexpr int PAW_GM_THREADS = 256, PAW_GM_PER_THREAD = 8;   // a CTA covers 2048 map cells of one word
/// recalculation (every 32 maintenance rounds): fixed-point sum of the word's occupancy map (:1304-1311)
__global__ void __launch_bounds__(PAW_GM_THREADS) pawcs_gmaint_sum(const PawArgs A) {
    GDict* gd = A.gd;
    const int g = blockIdx.y; // identity; every word is maintained exactly once whatever the dictionary order
    if(!(gd->weight[g] > 0.0f)) return;
    const int n = A.gW * A.gH;
    const float* m = A.gmap + (size_t)g * n;
    long long acc = 0;
    const int i0 = blockIdx.x * PAW_GM_THREADS * PAW_GM_PER_THREAD + threadIdx.x;
#pragma unroll
    for(int k = 0; k < PAW_GM_PER_THREAD; ++k) { const int i = i0 + k * PAW_GM_THREADS; if(i < n) acc += __double2ll_rn((double)m[i] * 4294967296.0); }
#pragma unroll
    for(int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, o);
    if((threadIdx.x & 31) == 0 && acc) atomicAdd(&gd->mapsum[g], (unsigned long long)acc);
}
/// one thread per word: new weight after a recalculation (below 1: the word dies and its map is zeroed), which words take part in
/// the update, weight *= 0.9 (:1312-1314). mflags: bit 0 zero the map, bit 1 update the map
__global__ void __launch_bounds__(PAW_MAXG) pawcs_gmaint_weights(const PawArgs A, int recalc, int update) {
    GDict* gd = A.gd;
    const int g = threadIdx.x;
    if(g >= A.NG) return;
    const bool live = gd->weight[g] > 0.0f;
    uint32_t flags = 0;
    if(recalc && live) {
        float w = (float)((double)(long long)gd->mapsum[g] / 4294967296.0);
        if(w < 1.0f) { w = 0.0f; flags |= 1u; }
        gd->weight[g] = w;
        if(w > 0.0f) flags |= 2u;
    } else if(live) flags |= 2u;
    if(update && (flags & 2u)) gd->weight[g] = __fmul_rn(gd->weight[g], 0.9f);
    gd->mflags[g] = flags; gd->mapsum[g] = 0ull;
}
/// accumulateProduct(map, -0.1, map, mask = nearest-downscaled ~dilate(lastFG)) followed by blur 3x3 (replicated border) (:1312-1315),
/// in one pass: the decayed neighbours are recomputed on the fly, the result goes to the scratch map
__global__ void __launch_bounds__(PAW_GM_THREADS) pawcs_gmaint_blur(const PawArgs A, int update) {
    const GDict* gd = A.gd;
    const int g = blockIdx.y, n = A.gW * A.gH;
    const uint32_t flags = gd->mflags[g];
    float* m = A.gmap + (size_t)g * n;
    float* tmp = A.gmap_tmp + (size_t)g * n;
    const int i0 = blockIdx.x * PAW_GM_THREADS * PAW_GM_PER_THREAD + threadIdx.x;
    if(flags & 1u) {
#pragma unroll
        for(int k = 0; k < PAW_GM_PER_THREAD; ++k) { const int i = i0 + k * PAW_GM_THREADS; if(i < n) m[i] = 0.0f; }
        return;
    }
    if(!(update && (flags & 2u))) return;
    auto decayed = [&](int cx, int cy) {
        const float v = m[(size_t)cy * A.gW + cx];
        const int x = cx * 2, y = cy * 2;
        return ((A.dilinv_bits[y * A.WW + (x >> 5)] >> (x & 31)) & 1u) ? __fadd_rn(v, __fmul_rn(v, -0.1f)) : v;
    };
#pragma unroll 1
    for(int k = 0; k < PAW_GM_PER_THREAD; ++k) {
        const int i = i0 + k * PAW_GM_THREADS;
        if(i >= n) break;
        const int cx = i % A.gW, cy = i / A.gW;
        const int xl = max(cx - 1, 0), xr = min(cx + 1, A.gW - 1);
        double s = 0.0;
#pragma unroll
        for(int d = -1; d <= 1; ++d) {
            const int yy = min(max(cy + d, 0), A.gH - 1);
            s += ((double)decayed(xl, yy) + (double)decayed(cx, yy)) + (double)decayed(xr, yy);
        }
        tmp[i] = (float)(s * (1.0 / 9.0));
    }
}
__global__ void __launch_bounds__(PAW_GM_THREADS) pawcs_gmaint_copy(const PawArgs A) {
    const int g = blockIdx.y, n = A.gW * A.gH;
    if(!(A.gd->mflags[g] & 2u)) return;
    float* m = A.gmap + (size_t)g * n;
    const float* tmp = A.gmap_tmp + (size_t)g * n;
    const int i0 = blockIdx.x * PAW_GM_THREADS * PAW_GM_PER_THREAD + threadIdx.x;
#pragma unroll
    for(int k = 0; k < PAW_GM_PER_THREAD; ++k) { const int i = i0 + k * PAW_GM_THREADS; if(i < n) m[i] = tmp[i]; }
}
__global__ void pawcs_gdict_bubble(const PawArgs A) { paw_gdict_bubble_pass(A); }
/// one bubble pass over the per-pixel global-word LUT (:1319-1334 / :411-428); `only_if_refresh`: last kernel of a conditional refresh:
/// runs only when a refresh just happened, and the last CTA to finish then bumps the epoch the refresh kernels consumed and clears the request
__global__ void __launch_bounds__(256) pawcs_glut_bubble(const PawArgs A, int only_if_refresh) {
    if(only_if_refresh && !A.gd->refresh_req) return;
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if(x < A.W && y < A.H && ((A.roi_bits[y * A.WW + (x >> 5)] >> (x & 31)) & 1u)) {
        const size_t pix = (size_t)y * A.Wp + x, cell = (size_t)(y >> 1) * A.gW + (x >> 1), msz = (size_t)A.gW * A.gH;
        uint32_t prev = A.glut[pix];
        float last = A.gmap[(size_t)prev * msz + cell];
        for(int i = 1; i < A.NG; ++i) {
            const uint32_t g = A.glut[(size_t)i * A.plane + pix];
            const float w = A.gmap[(size_t)g * msz + cell];
            if(w > last) { A.glut[(size_t)i * A.plane + pix] = (uchar)prev; A.glut[(size_t)(i - 1) * A.plane + pix] = (uchar)g; }
            else { last = w; prev = g; }
        }
    }
    if(only_if_refresh) {
        __syncthreads();
        if(threadIdx.x == 0 && threadIdx.y == 0) {
            __threadfence();
            if(atomicAdd(&A.gd->refresh_ticket, 1u) == gridDim.x * gridDim.y - 1u) {
                A.gd->refresh_ticket = 0u;
                A.ctl->refresh_epoch += 1; A.gd->refresh_req = PAW_REQ_NONE; A.gd->set_T_one = 0;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// frame-level analysis and tail (PAWCS.cpp:1462-1516)
// ------------------------------------------------------------------------------------------------------------
/// 8x8 area mean -> two EMAs -> masked fixed-point L1(LT,ST) (:1475-1478)
template<int CH>
__global__ void __launch_bounds__(128) pawcs_motion_kernel(const PawArgs A) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    long long acc = 0;
    if(i < A.dsW * A.dsH) {
        const int dx = i % A.dsW, dy = i / A.dsW;
        const float aLT = A.ctl->aLT, aST = A.ctl->aST;
        float vv[CH];
        if((A.W & 7) == 0 && (A.H & 7) == 0) { // OpenCV's integer-scale fast path: exact 8x8 mean
            uint32_t sum[CH];
#pragma unroll
            for(int c = 0; c < CH; ++c) sum[c] = 0;
            for(int r = 0; r < 8; ++r) {
                const uchar* p = A.img + (size_t)(dy * 8 + r) * A.ipitch + (size_t)dx * 8 * CH;
#pragma unroll
                for(int b = 0; b < 8; ++b)
#pragma unroll
                    for(int c = 0; c < CH; ++c) sum[c] += p[b * CH + c];
            }
#pragma unroll
            for(int c = 0; c < CH; ++c) vv[c] = fminf(fmaxf(rintf(__fmul_rn((float)sum[c], 1.0f / 64)), 0.f), 255.f);
        } else area_general_pixel<CH>(A.img, A.ipitch, A.W, A.H, A.dsW, A.dsH, dx, dy, vv);
        float t = 0.0f;
#pragma unroll
        for(int c = 0; c < CH; ++c) {
            const float v = vv[c];
            const size_t k = (size_t)i * CH + c;
            const float lt = __fadd_rn(__fmul_rn(v, aLT), __fmul_rn(A.dsLT[k], __fsub_rn(1.0f, aLT)));
            const float st = __fadd_rn(__fmul_rn(v, aST), __fmul_rn(A.dsST[k], __fsub_rn(1.0f, aST)));
            A.dsLT[k] = lt; A.dsST[k] = st;
            t = __fadd_rn(t, fabsf(__fsub_rn(lt, st)));
        }
        if(A.ds_roi[i]) acc = __double2ll_rn((double)t * 65536.0);
    }
#pragma unroll
    for(int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, o);
    if((threadIdx.x & 31) == 0 && acc) atomicAdd((unsigned long long*)&A.gd->motion_acc, (unsigned long long)acc);
}

/// getBackgroundImage / getBackgroundDescriptorsImage (:1525-1594): weighted mean over the local words, convertTo 8U / 16U
template<int CH>
__global__ void __launch_bounds__(256) pawcs_background_kernel(const PawArgs A, uchar* out_color, ushort* out_desc, uint32_t frame_off) {
    typedef typename Pack<CH>::Col Col;
    typedef typename Pack<CH>::Desc Desc;
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if(x >= A.W || y >= A.H) return;
    const size_t o = ((size_t)y * A.W + x) * CH;
    if(!((A.roi_bits[y * A.WW + (x >> 5)] >> (x & 31)) & 1u)) {
#pragma unroll
        for(int c = 0; c < CH; ++c) { if(out_color) out_color[o + c] = 0; if(out_desc) out_desc[o + c] = 0; }
        return;
    }
    const size_t pix = (size_t)y * A.Wp + x;
    const uint32_t frame = A.ctl->frame_idx - frame_off, woff = A.gd->weight_offset; // m_nFrameIdx: frame_off 0 inside a frame, 1 between frames
    const uint32_t wk = paw_wk(frame, woff);
    float tw = 0.0f, tc[CH], td[CH];
#pragma unroll
    for(int c = 0; c < CH; ++c) { tc[c] = 0.0f; td[c] = 0.0f; }
    for(int i = 0; i < A.NW; ++i) {
        const size_t at = (size_t)i * A.plane + pix;
        const float w = paw_weight(A.lw_key[at], wk);
        const typename PawRec<CH>::T rec = PAW_REC(A)[at];
        if(out_color) { const Col bc = PawRec<CH>::col(rec);
#pragma unroll
            for(int c = 0; c < CH; ++c) tc[c] = __fadd_rn(tc[c], __fmul_rn((float)col_get(bc, c), w)); }
        if(out_desc) { const Desc bd = PawRec<CH>::desc(rec);
#pragma unroll
            for(int c = 0; c < CH; ++c) td[c] = __fadd_rn(td[c], __fmul_rn((float)desc_get(bd, c), w)); }
        tw = __fadd_rn(tw, w);
    }
#pragma unroll
    for(int c = 0; c < CH; ++c) {
        if(out_color) out_color[o + c] = (uchar)fminf(fmaxf(rintf(__fdiv_rn(tc[c], tw)), 0.f), 255.f);
        if(out_desc) out_desc[o + c] = (ushort)fminf(fmaxf(rintf(__fdiv_rn(td[c], tw)), 0.f), 65535.f);
    }
}
/// model-vs-scene distances of the 500-frame check (:1483-1490): area-downsampled background image against the LT mean
template<int CH>
__global__ void __launch_bounds__(128) pawcs_model_dist_kernel(const PawArgs A) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    long long l1acc = 0, cdacc = 0;
    if(i < A.dsW * A.dsH && A.ds_roi[i] == 255) {
        const int dx = i % A.dsW, dy = i / A.dsW;
        float bg[CH], lt[CH];
        if((A.W & 7) == 0 && (A.H & 7) == 0) {
#pragma unroll
            for(int c = 0; c < CH; ++c) {
                uint32_t s = 0;
                for(int r = 0; r < 8; ++r) for(int b = 0; b < 8; ++b) s += A.bgimg[((size_t)(dy * 8 + r) * A.W + dx * 8 + b) * CH + c];
                bg[c] = fminf(fmaxf(rintf(__fmul_rn((float)s, 1.0f / 64)), 0.f), 255.f);
            }
        } else area_general_pixel<CH>(A.bgimg, (size_t)A.W * CH, A.W, A.H, A.dsW, A.dsH, dx, dy, bg);
#pragma unroll
        for(int c = 0; c < CH; ++c) lt[c] = A.dsLT[(size_t)i * CH + c];
        float t = 0.0f;
#pragma unroll
        for(int c = 0; c < CH; ++c) t = __fadd_rn(t, fabsf(__fsub_rn(lt[c], bg[c])));
        l1acc = __double2ll_rn((double)t * 65536.0);
        if(CH == 3) { // math.hpp:498-527 (float cdist)
            bool nonconst = false, nonnull = lt[0] != bg[0];
#pragma unroll
            for(int c = 1; c < CH; ++c) { nonconst |= (lt[c] != lt[c - 1]) || (bg[c] != bg[c - 1]); nonnull |= lt[c] != bg[c]; }
            if(nonconst && nonnull) {
                float cs = 0.0f, bs = 0.0f, mix = 0.0f;
#pragma unroll
                for(int c = 0; c < CH; ++c) { cs = __fadd_rn(cs, __fmul_rn(lt[c], lt[c])); bs = __fadd_rn(bs, __fmul_rn(bg[c], bg[c])); mix = __fadd_rn(mix, __fmul_rn(lt[c], bg[c])); }
                bs = __fadd_rn(bs, FLT_EPSILON);
                const float q = __fdiv_rn(__fmul_rn(mix, mix), bs);
                if(!(cs <= q)) cdacc = __double2ll_rn((double)__fsqrt_rn(__fsub_rn(cs, q)) * 65536.0);
            }
        }
    }
#pragma unroll
    for(int o = 16; o > 0; o >>= 1) { l1acc += __shfl_xor_sync(0xFFFFFFFFu, l1acc, o); cdacc += __shfl_xor_sync(0xFFFFFFFFu, cdacc, o); }
    if((threadIdx.x & 31) == 0) {
        if(l1acc) atomicAdd((unsigned long long*)&A.gd->model_l1_acc, (unsigned long long)l1acc);
        if(cdacc) atomicAdd((unsigned long long*)&A.gd->model_cd_acc, (unsigned long long)cdacc);
    }
}

/// tail, part 1 (1 CTA, 256 threads): LUT adaptation (:1462-1473), auto-reset enable, moving-camera decision (:1479-1500).
/// `check_model`: this frame is a multiple of the bootstrap window and the model distances were accumulated.
__device__ __forceinline__ void paw_tail2(const PawArgs& A);
__global__ void __launch_bounds__(256) pawcs_tail1_kernel(const PawArgs A, int check_model, int then_tail2) {
    FrameCtl* ctl = A.ctl; GDict* gd = A.gd;
    __shared__ int s_dir;
    const int t = threadIdx.x;
    if(t == 0) {
        const float ratio = __fdiv_rn((float)(ctl->roi_count - gd->flat_count), (float)ctl->roi_count);
        const float last = gd->last_nonflat_ratio;
        s_dir = (ratio < 0.1f && last < 0.1f) ? -1 : (ratio > 0.5f && last > 0.5f) ? 1 : 0;
        gd->last_nonflat_ratio = ratio; gd->flat_count = 0;
        gd->refresh_req = PAW_REQ_NONE; gd->set_T_one = 0;
        const float l1ratio = __fdiv_rn((float)((double)gd->motion_acc / 65536.0), (float)gd->ds_roi_count);
        if(!ctl->auto_reset && l1ratio >= 90.0f) ctl->auto_reset = 1;
        gd->tail_gate = (ctl->auto_reset || gd->moving_camera) ? 1u : 0u; // evaluated once, before the mode can change (:1481)
        if(gd->tail_gate && check_model) {
            const float ml1 = __fdiv_rn((float)((double)gd->model_l1_acc / 65536.0), (float)gd->ds_roi_count);
            const float mcd = __fdiv_rn((float)((double)gd->model_cd_acc / 65536.0), (float)gd->ds_roi_count);
            if(gd->moving_camera && ml1 < 11.0f && mcd < 1.0f) {
                gd->weight_offset = PAW_WEIGHT_OFFSET; gd->moving_camera = 0;
                gd->refresh_req = PAW_REQ_REFRESH; gd->refresh_base_occ = 1; gd->refresh_decr = 1.0f; gd->refresh_force = 1;
            } else if(gd->boot && !gd->moving_camera && (ml1 >= 45.0f || mcd >= 4.0f)) {
                gd->weight_offset = 5; gd->moving_camera = 1;
                gd->refresh_req = PAW_REQ_REFRESH; gd->refresh_base_occ = 1; gd->refresh_decr = 1.0f; gd->refresh_force = 1;
            }
        }
        gd->model_l1_acc = 0; gd->model_cd_acc = 0;
    }
    __syncthreads();
    const int dir = s_dir;
    if(dir < 0) {
        const float lo = fminf(fmaxf(rintf(__fdiv_rn(__fadd_rn((float)A.lbsp_off, __fmul_rn((float)t, A.rel)), 4.0f)), 0.f), 255.f);
        if((float)A.lut[t] > lo) A.lut[t] -= 1;
    } else if(dir > 0) {
        const float hi = fminf(fmaxf(rintf(__fadd_rn((float)A.lbsp_off, __fmul_rn(255.0f, A.rel))), 0.f), 255.f);
        if((float)A.lut[t] < hi) A.lut[t] += 1;
    }
    if(then_tail2 && t == 0) paw_tail2(A);
}
/// tail, part 2 (1 thread): reset logic (:1501-1515), next-frame factors; follows part 1 in the same launch (`then_tail2`) unless the
/// 500-frame model check may put a refresh between the two
__device__ __forceinline__ void paw_tail2(const PawArgs& A) {
    FrameCtl* ctl = A.ctl; GDict* gd = A.gd;
    gd->refresh_req = PAW_REQ_NONE;
    const float l1ratio = __fdiv_rn((float)((double)gd->motion_acc / 65536.0), (float)gd->ds_roi_count);
    gd->motion_acc = 0;
    if(gd->tail_gate) {
        if(ctl->frames_since_reset > PAW_BOOTSTRAP * 2u) ctl->auto_reset = 0;
        else if(l1ratio >= 45.0f && ctl->cooldown == 0) {
            ctl->frames_since_reset = 0;
            gd->refresh_req = PAW_REQ_REFRESH; gd->refresh_base_occ = gd->weight_offset / 8u; gd->refresh_decr = 0.0f; gd->refresh_force = 1;
            ctl->cooldown = gd->nST;
            gd->set_T_one = 1;
        } else if(!gd->boot) ctl->frames_since_reset += 1;
    }
    if(ctl->cooldown > 0) ctl->cooldown -= 1;
    ctl->wl_count = 0; gd->wlB_count = 0; // this frame's work-lists are consumed
    // next frame
    const uint32_t f = ctl->frame_idx + 1;
    ctl->frame_idx = f;
    const bool boot = f <= PAW_BOOTSTRAP;
    gd->boot = boot;
    const uint32_t nLT = boot ? (uint32_t)A.avg_samples / 2u : (uint32_t)A.avg_samples, nST = nLT / 4u;
    gd->nST = nST;
    ctl->aLT = __fdiv_rn(1.0f, (float)min(f, nLT));
    ctl->aST = __fdiv_rn(1.0f, (float)min(f, nST));
}
__global__ void pawcs_tail2_kernel(const PawArgs A) { paw_tail2(A); }

// ------------------------------------------------------------------------------------------------------------
// refreshModel (PAWCS.cpp:107-429); runs only when gd->refresh_req is set (the tail decides on the device)
// ------------------------------------------------------------------------------------------------------------
template<int CH>
__global__ void __launch_bounds__(256) pawcs_refresh_local(const PawArgs A, uint32_t frame_off) {
    typedef typename Pack<CH>::Col Col;
    typedef typename Pack<CH>::Desc Desc;
    const GDict* gd = A.gd;
    if(!gd->refresh_req) return;
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if(x >= A.W || y >= A.H) return;
    const size_t pix = (size_t)y * A.Wp + x;
    if(gd->set_T_one) A.maps[pix * 2].x = 1.0f;
    const int wi = y * A.WW + (x >> 5);
    if(!((A.roi_bits[wi] >> (x & 31)) & 1u)) return;
    const bool force = gd->refresh_force != 0;
    if(!force && ((A.dil_bits[wi] >> (x & 31)) & 1u)) return;
    const uint32_t frame = A.ctl->frame_idx - frame_off; // m_nFrameIdx (0 at initialisation)
    const uint32_t epoch = A.ctl->refresh_epoch, woff = gd->weight_offset, base_occ = gd->refresh_base_occ;
    const float decr = gd->refresh_decr;
    const uint32_t pixid = (uint32_t)(y * A.W + x);
    const float R = A.maps[pix * 2].y;
    const bool unst = (A.unstable_bits[wi] >> (x & 31)) & 1u;
    const uint32_t cbase = (uint32_t)__fmul_rn(__fsqrt_rn(R), (float)A.min_color);
    const uint32_t thrC = CH == 1 ? cbase / 2u : cbase * 3u;
    const uint32_t dbase = (1u << (uint32_t)floorf(__fadd_rn(R, 0.5f))) + (uint32_t)A.desc_off + (unst ? (uint32_t)A.desc_off : 0u);
    const uint32_t thrD = CH == 1 ? dbase : dbase * 3u;
    const int NW = A.NW;
    // occurrence == 0 && last == 0 && first == 1 marks a word that does not exist yet (initialisation only)
    auto valid = [&](int i) { const size_t at = (size_t)i * A.plane + pix; return !(PawRec<CH>::load_first(PAW_REC(A) + at) == 1u && A.lw_key[at].y == 1u); }; // first == 1 && last == 0
    const uint32_t wk = paw_wk(frame, woff);
    auto weight = [&](int i) { const size_t at = (size_t)i * A.plane + pix; return paw_weight(A.lw_key[at], wk); };
    if(decr > 0.0f)
        for(int i = 0; i < NW; ++i) { const size_t at = (size_t)i * A.plane + pix; if(valid(i)) { const uint32_t o = A.lw_key[at].x; A.lw_key[at].x = o - (uint32_t)__fmul_rn(decr, (float)o); } }
    uint32_t site = 0;
    uint4 rnd = make_uint4(0, 0, 0, 0);
    auto draw = [&]() { if((site & 3u) == 0u) rnd = philox_block(A.seed, epoch, pixid, site >> 2, DOM_REFRESH);
                        const uint32_t k = site & 3u; ++site; return k == 0 ? rnd.x : k == 1 ? rnd.y : k == 2 ? rnd.z : rnd.w; };
    for(int it = 0; it < 98; ++it) {
        int sx, sy;
        sample_pos_7x7(draw(), sx, sy, x, y, A.W, A.H);
        if(!force && ((A.dil_bits[sy * A.WW + (sx >> 5)] >> (sx & 31)) & 1u)) continue;
        const size_t sp = (size_t)sy * A.Wp + sx;
        const Col scol = ((const Col*)A.last_color)[sp];
        const Desc sdesc = ((const Desc*)A.last_desc)[sp];
        int i;
        for(i = 0; i < NW; ++i) {
            if(!valid(i)) continue;
            const size_t at = (size_t)i * A.plane + pix;
            uint32_t l1, cd;
            const typename PawRec<CH>::T r = PAW_REC(A)[at];
            if(paw_color_dist<CH>(col_as_u32(scol), col_as_u32(PawRec<CH>::col(r)), l1, cd) <= thrC && paw_hdist(sdesc, PawRec<CH>::desc(r)) <= thrD) {
                A.lw_key[at] = make_uint2(A.lw_key[at].x + 1u, PawRec<CH>::first(r) + frame); break; // last = frame
            }
        }
        if(i == NW) {
            i = NW - 1;
            const size_t at = (size_t)i * A.plane + pix;
            PAW_REC(A)[at] = PawRec<CH>::make(scol, sdesc, frame);
            A.lw_key[at] = make_uint2(base_occ, frame * 2u);
        }
        while(i > 0 && (!valid(i - 1) || weight(i) > weight(i - 1))) { paw_swap<CH>(A, pix, i); --i; }
    }
    for(int i = 1; i < NW; ++i) { // random resampling of the words still missing (:322-339)
        if(valid(i)) continue;
        const size_t at = (size_t)i * A.plane + pix;
        const uint32_t r = draw() % (uint32_t)i;
        const size_t ar = (size_t)r * A.plane + pix;
        const uint32_t d2 = draw();
        const int off = CH == 1 ? (int)(d2 % (thrC + 1u)) - (int)thrC / 2 : (int)(d2 % (thrC / 3u + 1u)) - (int)(thrC / 6u);
        const typename PawRec<CH>::T rr = PAW_REC(A)[ar];
        const Col rc = PawRec<CH>::col(rr);
        Col nc;
        if constexpr (CH == 1) nc = (uchar)min(max((int)rc + off, 0), 255);
        else nc = (uint32_t)min(max((int)(rc & 0xFFu) + off, 0), 255) | ((uint32_t)min(max((int)((rc >> 8) & 0xFFu) + off, 0), 255) << 8)
                | ((uint32_t)min(max((int)((rc >> 16) & 0xFFu) + off, 0), 255) << 16);
        PAW_REC(A)[at] = PawRec<CH>::make(nc, PawRec<CH>::desc(rr), frame);
        const uint32_t o = (uint32_t)__fmul_rn((float)A.lw_key[ar].x, __fdiv_rn((float)(NW - i), (float)NW));
        A.lw_key[at] = make_uint2(max(o, 1u), frame * 2u);
    }
}
/// global resampling (:342-408): sequential by nature (<= ~4*NG pixels); thread 0 decides, the CTA zeroes maps
template<int CH>
__global__ void __launch_bounds__(1024) pawcs_refresh_global(const PawArgs A, uint32_t frame_off) {
    typedef typename Pack<CH>::Col Col;
    typedef typename Pack<CH>::Desc Desc;
    GDict* gd = A.gd;
    if(!gd->refresh_req) return;
    __shared__ int s_zero_g;
    const bool force = gd->refresh_force != 0;
    const uint32_t frame = A.ctl->frame_idx - frame_off, woff = gd->weight_offset;
    const size_t npx = (size_t)A.W * A.H, msz = (size_t)A.gW * A.gH;
    const int NG = A.NG;
    size_t incr = max(npx / (size_t)NG, (size_t)1);
    for(int pass = 0; pass < 2; ++pass) {
        for(size_t p = 0; p < npx; p += incr) {
            const int x = (int)(p % A.W), y = (int)(p / A.W);
            const int wi = y * A.WW + (x >> 5);
            if(!((A.roi_bits[wi] >> (x & 31)) & 1u)) continue;               // uniform across the CTA
            if(!force && ((A.dil_bits[wi] >> (x & 31)) & 1u)) continue;
            const size_t pix = (size_t)y * A.Wp + x;
            int i = 0;
            if(threadIdx.x == 0) {
                s_zero_g = -1;
                const float R = A.maps[pix * 2].y;
                const bool unst = (A.unstable_bits[wi] >> (x & 31)) & 1u;
                const uint32_t cbase = (uint32_t)__fmul_rn(__fsqrt_rn(R), (float)A.min_color);
                const uint32_t thrC = CH == 1 ? cbase / 2u : cbase * 3u;
                const uint32_t dbase = (1u << (uint32_t)floorf(__fadd_rn(R, 0.5f))) + (uint32_t)A.desc_off + (unst ? (uint32_t)A.desc_off : 0u);
                const uint32_t thrD = CH == 1 ? dbase : dbase * 3u;
                const typename PawRec<CH>::T r0 = PAW_REC(A)[pix];
                const Col bc = PawRec<CH>::col(r0); const Desc bd = PawRec<CH>::desc(r0);
                const uint32_t bits = paw_bits(bd);
                bool found_uninit = false;
                for(i = 0; i < NG; ++i) {
                    const int g = gd->dict[i];
                    if(g < 0) { found_uninit = true; continue; }
                    const uint32_t gb = gd->bits[g];
                    uint32_t l1, cd;
                    if((bits > gb ? bits - gb : gb - bits) <= thrD / 4u && paw_color_dist<CH>(col_as_u32(bc), gd->color[g], l1, cd) <= thrC) break;
                }
                if(i == NG) {
                    i = NG - 1;
                    const int g = found_uninit ? (int)gd->created++ : gd->dict[i];
                    gd->color[g] = col_as_u32(bc);
                    if constexpr (CH == 1) gd->desc[g] = make_uint2(bd, 0); else gd->desc[g] = bd;
                    gd->bits[g] = bits; gd->weight[g] = 0.0f; gd->dict[i] = g;
                    s_zero_g = g;
                }
            }
            __syncthreads();
            const int zg = s_zero_g;
            if(zg >= 0) { float* m = A.gmap + (size_t)zg * msz; for(size_t k = threadIdx.x; k < msz; k += blockDim.x) m[k] = 0.0f; }
            __syncthreads();
            if(threadIdx.x == 0) {
                const int g = gd->dict[i];
                const float bw = paw_weight(A.lw_key[pix], paw_wk(frame, woff));
                float* cw = A.gmap + (size_t)g * msz + (size_t)(y >> 1) * A.gW + (x >> 1);
                if(*cw < bw) { gd->weight[g] = __fadd_rn(gd->weight[g], bw); *cw = __fadd_rn(*cw, bw); }
                while(i > 0 && (gd->dict[i - 1] < 0 || gd->weight[gd->dict[i]] > gd->weight[gd->dict[i - 1]])) { const int t = gd->dict[i]; gd->dict[i] = gd->dict[i - 1]; gd->dict[i - 1] = t; --i; }
            }
            __syncthreads();
        }
        incr = max(incr / 3, (size_t)1);
    }
    for(int i = 0; i < NG; ++i) { // :397-408 (initialisation only)
        if(gd->dict[i] >= 0) continue;
        __syncthreads();
        if(threadIdx.x == 0) { const int g = (int)gd->created++; gd->color[g] = 0; gd->desc[g] = make_uint2(0, 0); gd->bits[g] = 0; gd->weight[g] = 0.0f; gd->dict[i] = g; s_zero_g = g; }
        __syncthreads();
        float* m = A.gmap + (size_t)s_zero_g * msz;
        for(size_t k = threadIdx.x; k < msz; k += blockDim.x) m[k] = 0.0f;
        __syncthreads();
    }
}
} // namespace lvb
