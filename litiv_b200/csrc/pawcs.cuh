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
// HBM layout (per stream; Wp = W rounded up to 32): local words are five sample-major SoA planes [NW][H][Wp]
// (first u32, last u32, occurrences u32, colour u32 B,G,R,0, descriptors uint2): every frame every pixel re-weights all
// NW words (12 B each, three coalesced 128-byte requests per warp and word), colour/descriptor are touched only while
// the weight sum is below its threshold. Global words: GDict (small arrays) + occupancy maps [NG][H/2][W/2] f32 +
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
};

struct PawArgs {
    int W, H, Wp, WW, NW, NG, gW, gH;
    size_t plane;
    const uchar* img; size_t ipitch;
    uint32_t* lw_first; uint32_t* lw_last; uint32_t* lw_occ; void* lw_color; void* lw_desc;
    uchar* glut; float* gmap; float* gmap_tmp; GDict* gd;
    float4* maps; float2* fin; void* last_color; void* last_desc;
    const uint32_t* roi_bits; const uint32_t* roi255_bits;
    uint32_t* raw_bits; uint32_t* unstable_bits; const uint32_t* blinks_bits; const uint32_t* lastfg_bits;
    uint32_t* illum_bits; uint32_t* did_bits; const uint32_t* dil_bits; const uint32_t* dilinv_bits;
    uint32_t* intent_bits; uint4* intents; size_t bitplane;
    uint32_t* gop_bits; float* gop_w; uchar* gop_g;
    uchar* lut; FrameCtl* ctl;
    uint64_t seed; uint32_t lr_fixed; int min_color, desc_off;
    int use_tma, collect_stats;
    float rel; int lbsp_off, avg_samples, dsW, dsH;
    const uchar* ds_roi; float* dsLT; float* dsST; uchar* bgimg; // frame-level analysis
};

__device__ __forceinline__ float paw_weight(uint32_t first, uint32_t last, uint32_t occ, uint32_t frame, uint32_t off) { // PAWCS.cpp:1596-1598
    return __fdiv_rn((float)occ, (float)((last - first) + (frame - last) * 2u + off));
}
__device__ __forceinline__ uint32_t paw_hdist(const uint2& a, const uint2& b) { return __popc(a.x ^ b.x) + __popc((a.y ^ b.y) & 0xFFFFu); }
__device__ __forceinline__ uint32_t paw_hdist(const ushort& a, const ushort& b) { return __popc((uint32_t)(a ^ b)); }
__device__ __forceinline__ uint32_t paw_bits(const uint2& a) { return __popc(a.x) + __popc(a.y & 0xFFFFu); }
__device__ __forceinline__ uint32_t paw_bits(const ushort& a) { return __popc((uint32_t)a); }

/// colour distances of math.hpp: L1dist (u8-wrapping for 3 channels, Q1), cdist :474-496, cmixdist :596-605
template<int CH>
__device__ __forceinline__ uint32_t paw_color_dist(uint32_t cur, uint32_t bg, uint32_t& l1, uint32_t& cd) {
    if(CH == 1) { const uint32_t a = cur & 0xFFu, b = bg & 0xFFu; l1 = a > b ? a - b : b - a; cd = 0; return l1; }
    const uint32_t c0 = cur & 0xFFu, c1 = (cur >> 8) & 0xFFu, c2 = (cur >> 16) & 0xFFu;
    const uint32_t b0 = bg & 0xFFu, b1 = (bg >> 8) & 0xFFu, b2 = (bg >> 16) & 0xFFu;
    l1 = (__usad(c0, b0, 0u) + __usad(c1, b1, 0u) + __usad(c2, b2, 0u)) & 0xFFu;
    const bool nonconst = (c1 != c0) || (b1 != b0) || (c2 != c1) || (b2 != b1);
    const bool nonnull = (c0 != b0) || (c1 != b1) || (c2 != b2);
    cd = 0;
    if(nonconst && nonnull) {
        const uint32_t cs = c0 * c0 + c1 * c1 + c2 * c2, bs = b0 * b0 + b1 * b1 + b2 * b2, mix = c0 * b0 + c1 * b1 + c2 * b2;
        // floor(mix^2 / max(bs,1)) exactly, without a 64-bit or double division: q <= cs < 2^18 (Cauchy-Schwarz), so a single-
        // precision estimate (relative error < 2^-21) is off by at most one; the 64-bit remainder fixes it
        const unsigned long long m2 = (unsigned long long)mix * mix;
        const uint32_t d = max(bs, 1u);
        uint32_t q = (uint32_t)__fdividef(__fmul_rn((float)mix, (float)mix), (float)d);
        long long r = (long long)m2 - (long long)((unsigned long long)q * d);
        while(r < 0) { --q; r += d; }
        while(r >= (long long)d) { ++q; r -= d; }
        cd = (uint32_t)__fsqrt_rn((float)(cs - q));
    }
    return (l1 >> 1) + cd * 4u;
}

template<int CH> struct PawPlanes {
    typedef typename Pack<CH>::Col Col;
    typedef typename Pack<CH>::Desc Desc;
};
__device__ __forceinline__ uint32_t col_as_u32(const uint32_t& v) { return v; }
__device__ __forceinline__ uint32_t col_as_u32(const uchar& v) { return v; }

/// exchange dictionary positions i and i-1 of one pixel (all five planes)
template<int CH>
__device__ __forceinline__ void paw_swap(const PawArgs& A, size_t pix, int i) {
    typedef typename Pack<CH>::Col Col;
    typedef typename Pack<CH>::Desc Desc;
    const size_t a = (size_t)i * A.plane + pix, b = a - A.plane;
    const uint32_t f = A.lw_first[a], l = A.lw_last[a], o = A.lw_occ[a];
    const Col c = ((Col*)A.lw_color)[a]; const Desc d = ((Desc*)A.lw_desc)[a];
    A.lw_first[a] = A.lw_first[b]; A.lw_last[a] = A.lw_last[b]; A.lw_occ[a] = A.lw_occ[b];
    ((Col*)A.lw_color)[a] = ((Col*)A.lw_color)[b]; ((Desc*)A.lw_desc)[a] = ((Desc*)A.lw_desc)[b];
    A.lw_first[b] = f; A.lw_last[b] = l; A.lw_occ[b] = o; ((Col*)A.lw_color)[b] = c; ((Desc*)A.lw_desc)[b] = d;
}

/// search the pixel's sorted global-word LUT (PAWCS.cpp:1073-1079 / :1119-1125); returns the identity or -1
template<int CH>
__device__ __forceinline__ int paw_find_gword(const PawArgs& A, const uint32_t* s_gbits, const uint32_t* s_gcolor, size_t pix, uint32_t cur_pack, uint32_t bits, uint32_t thrC, uint32_t thrD) {
    for(int gi = 0; gi < A.NG; ++gi) {
        const int g = A.glut[(size_t)gi * A.plane + pix];
        const uint32_t gb = s_gbits[g];
        if((bits > gb ? bits - gb : gb - bits) <= thrD / 4u) {
            uint32_t l1, cd;
            if(paw_color_dist<CH>(cur_pack, s_gcolor[g], l1, cd) <= thrC) return g;
        }
    }
    return -1;
}

#ifndef PAW_CHUNK
#define PAW_CHUNK 8   // words of the bubble pass whose counters are in flight together
#endif
#ifndef PAW_MIN_BLOCKS
#define PAW_MIN_BLOCKS 4   // 64 registers with some spills, 4 CTAs per SM: the kernel is latency bound (2 / 3 / 4 / 5 / 6 -> 2.90 / 2.63 / 2.52 / 2.67 / 2.58 ms)
#endif
template<int CH, bool T7>
__global__ void __launch_bounds__(TILE_W * TILE_H, PAW_MIN_BLOCKS)
pawcs_phaseA(const PawArgs A, const __grid_constant__ CUtensorMap tmap) {
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
    uint32_t f0 = 0, l0 = 0, o0 = 0;
    if(active) {
        m0 = A.maps[pix * 2]; m1 = A.maps[pix * 2 + 1]; fin = A.fin[pix];
        f0 = A.lw_first[pix]; l0 = A.lw_last[pix]; o0 = A.lw_occ[pix];
    }
    stage_tile_wait(&s_bar, A.use_tma);

    bool seg = false, unstable_new = false, did = false, has_intent = false, has_gop = false, flat = false;
    uint32_t scanned = 0;
    int intent_row = 0;
    if(active) {
        const FrameCtl* ctl = A.ctl; const GDict* gd = A.gd;
        const float aLT = ctl->aLT, aST = ctl->aST;
        const uint32_t frame = ctl->frame_idx, cooldown = ctl->cooldown;
        const uint32_t woff = gd->weight_offset; const bool boot = gd->boot != 0, moving = gd->moving_camera != 0;
        const uint32_t colorRange = CH == 1 ? 255u : 765u, descRange = CH == 1 ? 16u : 48u, flatK = CH == 1 ? 2u : 4u;
        float T = m0.x, R = m0.y, V = m0.z, DminLT = m1.x, DminST = m1.y, rawLT = m1.z, rawST = m1.w;
        const bool unst = (w_unst & lane_bit) != 0, blink = (w_blink & lane_bit) != 0, lastfg = (w_lastfg & lane_bit) != 0;
        const bool border = !(w_roi255 & lane_bit);
        uint32_t illum_cur = (w_illum & lane_bit) ? 1u : 0u;

        const int sy = threadIdx.y + HALO;
        Lookup16 L[CH];
        uint32_t cur[CH], intra[CH];
        {
            const Window5<CH> Wn = lbsp_window_smem<CH>(s_tile, PITCH, sy, tile_shift(CH) + (int)threadIdx.x * CH);
#pragma unroll
            for(int c = 0; c < CH; ++c) {
                L[c] = lbsp_lookup_window<CH>(Wn, c);
                cur[c] = win_center<CH>(Wn, c);
                intra[c] = lbsp_threshold<T7>(L[c], cur[c], s_lut[cur[c]]);
            }
        }
        Col cur_pack; Desc intra_pack;
        if constexpr (CH == 1) { cur_pack = (uchar)cur[0]; intra_pack = (ushort)intra[0]; }
        else { cur_pack = cur[0] | (cur[1] << 8) | (cur[2] << 16); intra_pack = make_uint2(intra[0] | (intra[1] << 16), intra[2]); }
        const uint32_t cur32 = col_as_u32(cur_pack);
        const uint32_t bits = paw_bits(intra_pack);
        flat = bits < flatK;
        const uint32_t occ_incr = (1u + cooldown) << ((flat || boot) ? 1 : 0);
        const uint32_t rate = A.lr_fixed ? A.lr_fixed : (flat ? (uint32_t)ceilf(__fadd_rn(T, 1.0f)) / 2u : (uint32_t)ceilf(T)); // :987-989
        // thresholds (:993-994)
        const uint32_t cbase = (uint32_t)__fmul_rn(__fsqrt_rn(R), (float)A.min_color);
        const uint32_t thrC = CH == 1 ? cbase / 2u : cbase * 3u;
        const uint32_t dbase = (1u << (uint32_t)floorf(__fadd_rn(R, 0.5f))) + (uint32_t)A.desc_off + (unst ? (uint32_t)A.desc_off : 0u);
        const uint32_t thrD = CH == 1 ? dbase : dbase * 3u;
        const float wthr = __fdiv_rn(paw_weight(f0, l0, o0, frame, woff), __fmul_rn(R, 2.0f)); // :967-968
        const uint32_t pixid = (uint32_t)(y * A.W + x);

        // local dictionary: scan while the weight sum is below its threshold (:1002-1053), bubble pass over all words (:1044-1065)
        float sum = 0.0f, last_w = FLT_MAX;
        uint32_t minColor = colorRange, minDesc = descRange;
        int i = 0;
        uint32_t wf = f0, wl = l0, wo = o0;
        // software-pipelined scan: word i+1 (colour, descriptor, counters) is fetched while word i is tested. A swap at step i
        // exchanges positions i and i-1 only, so the prefetched word is still the one at position i+1.
        Col nbc = ((const Col*)A.lw_color)[pix];
        Desc nbd = ((const Desc*)A.lw_desc)[pix];
        uint32_t nwf = f0, nwl = l0, nwo = o0;
        for(; i < A.NW && sum < wthr; ++i) {
            const size_t at = (size_t)i * A.plane + pix;
            const Col bc = nbc;
            const Desc bd = nbd;
            wf = nwf; wl = nwl; wo = nwo;
            if(i + 1 < A.NW) {
                const size_t an = at + A.plane;
                nbc = ((const Col*)A.lw_color)[an]; nbd = ((const Desc*)A.lw_desc)[an];
                nwf = A.lw_first[an]; nwl = A.lw_last[an]; nwo = A.lw_occ[an];
            }
            const float w = paw_weight(wf, wl, wo, frame, woff);
            ++scanned;
            uint32_t l1, cd;
            const uint32_t mix = paw_color_dist<CH>(cur32, col_as_u32(bc), l1, cd);
            const uint32_t ihd = paw_hdist(intra_pack, bd);
            uint32_t ehd = 0;
#pragma unroll
            for(int c = 0; c < CH; ++c) {
                const uint32_t b = col_get(bc, c);
                ehd += __popc(lbsp_threshold<T7>(L[c], b, s_lut[b]) ^ desc_get(bd, c));
            }
            const uint32_t dd = (ihd + ehd) >> 1;
            if((!unst || flat || border) && mix <= thrC && l1 >= thrC / 2u && ihd <= thrD / 2u) { // illumination update (:1014-1030)
                const uint32_t mod = illum_cur ? (rate / 2u + 1u) : rate;
                if((philox_draw(A.seed, frame, pixid, 4u + (uint32_t)i, DOM_PAWCS_A) % mod) == 0u) {
                    ((Col*)A.lw_color)[at] = cur_pack; ((Desc*)A.lw_desc)[at] = intra_pack;
                    did = true; illum_cur = 2u;
                }
            }
            if(dd <= thrD && mix <= thrC) {
                sum = __fadd_rn(sum, w);
                A.lw_last[at] = frame;
                if((!lastfg || moving) && w < 1.0f) A.lw_occ[at] = wo + occ_incr;
                minColor = min(minColor, mix); minDesc = min(minDesc, dd);
            }
            if(w > last_w) paw_swap<CH>(A, pix, i); else last_w = w;
        }
        // the bubble pass continues over the rest of the dictionary (:1054-1065): only the counters are needed, CHUNK words
        // in flight at once (a swap exchanges positions i and i-1 in memory; words already in registers are at positions > i)
        constexpr int CHUNK = PAW_CHUNK;
        for(; i < A.NW; i += CHUNK) {
            uint32_t cf[CHUNK], cl[CHUNK], co[CHUNK];
#pragma unroll
            for(int k = 0; k < CHUNK; ++k) {
                if(i + k < A.NW) { const size_t at = (size_t)(i + k) * A.plane + pix; cf[k] = A.lw_first[at]; cl[k] = A.lw_last[at]; co[k] = A.lw_occ[at]; }
                else { cf[k] = 0; cl[k] = 0; co[k] = 0; }
            }
#pragma unroll
            for(int k = 0; k < CHUNK; ++k) {
                if(i + k < A.NW) {
                    const float w = paw_weight(cf[k], cl[k], co[k], frame, woff);
                    if(w > last_w) paw_swap<CH>(A, pix, i + k); else last_w = w;
                }
            }
        }

        const uint4 rnd = philox_block(A.seed, frame, pixid, 0, DOM_PAWCS_A);
        const float oneLT = __fsub_rn(1.0f, aLT), oneST = __fsub_rn(1.0f, aST);
        const float baseMin = fmaxf(__fdiv_rn((float)minColor, (float)colorRange), __fdiv_rn((float)minDesc, (float)descRange));
        const size_t cell = (size_t)(y >> 1) * A.gW + (x >> 1);
        if(sum >= wthr || border) { // background (:1070-1106)
            DminLT = __fadd_rn(__fmul_rn(DminLT, oneLT), __fmul_rn(baseMin, aLT));
            DminST = __fadd_rn(__fmul_rn(DminST, oneST), __fmul_rn(baseMin, aST));
            rawLT = __fmul_rn(rawLT, oneLT); rawST = __fmul_rn(rawST, oneST);
            if((rnd.x % rate) == 0u) {
                const int g = paw_find_gword<CH>(A, s_gbits, s_gcolor, pix, cur32, bits, thrC, thrD);
                const uint32_t rep = rate >= 0x40000000u ? rnd.y : rnd.y % (rate * 2u);
                if(g >= 0 || rep == 0u) {
                    A.gop_g[pix] = g >= 0 ? (uchar)g : (uchar)0xFE; A.gop_w[pix] = sum; has_gop = true;
                    if(g < 0) atomicMin(&A.gd->rep_winner, pixid);
                }
            }
        } else { // foreground (:1107-1155)
            const float nmin = fmaxf(baseMin, __fdiv_rn(__fsub_rn(wthr, sum), wthr));
            DminLT = __fadd_rn(__fmul_rn(DminLT, oneLT), __fmul_rn(nmin, aLT));
            DminST = __fadd_rn(__fmul_rn(DminST, oneST), __fmul_rn(nmin, aST));
            rawLT = __fadd_rn(__fmul_rn(rawLT, oneLT), aLT); rawST = __fadd_rn(__fmul_rn(rawST, oneST), aST);
            if(flat || (rnd.x % rate) == 0u) {
                const int g = paw_find_gword<CH>(A, s_gbits, s_gcolor, pix, cur32, bits, thrC, thrD);
                if(g < 0) seg = true;
                else if(__fadd_rn(sum, __fdiv_rn(A.gmap[(size_t)g * A.gW * A.gH + cell], flat ? 2.0f : 4.0f)) < wthr) seg = true;
            } else seg = true;
            if(sum < __fdiv_rn(1.0f, (float)woff)) { // new local word over the last one (:1142-1153)
                const size_t at = (size_t)(A.NW - 1) * A.plane + pix;
                ((Col*)A.lw_color)[at] = cur_pack; ((Desc*)A.lw_desc)[at] = intra_pack;
                A.lw_occ[at] = occ_incr; A.lw_first[at] = frame; A.lw_last[at] = frame;
            }
        }
        // neighbour dictionary update, queued (:1164-1247)
        if((!seg && (rnd.z % rate) == 0u) || border || moving) {
            int dx, dy;
            neighbor_offset(!(flat || border || moving), rnd.w, dx, dy);
            const int nx = clampi(x + dx, 2, A.W - 3), ny = clampi(y + dy, 2, A.H - 3);
            if((A.roi_bits[ny * A.WW + (nx >> 5)] >> (nx & 31)) & 1u) {
                A.intents[pix] = make_uint4((uint32_t)((ny - y + 2) * 5 + (nx - x + 2)) | (thrD << 8), thrC, __float_as_uint(wthr), rate);
                has_intent = true; intent_row = ny - y + 2;
            }
        }
        // feedback (:1252-1269)
        unstable_new = (R > 3.0f) || (__fsub_rn(rawLT, fin.x) > 0.1f) || (__fsub_rn(rawST, fin.y) > 0.1f);
        const float dmin = fminf(DminLT, DminST), dmax = fmaxf(DminLT, DminST);
        if(lastfg || (dmin < 0.1f && seg)) T = fminf(__fadd_rn(T, __fdiv_rn(0.5f, __fmul_rn(dmax, V))), 256.0f);
        else T = fmaxf(__fsub_rn(T, __fdiv_rn(__fmul_rn(0.25f, V), dmax)), 1.0f);
        if(dmax > 0.1f && blink) V = __fadd_rn(V, boot ? 2.0f : 1.0f);
        else V = fmaxf(__fsub_rn(V, __fmul_rn(0.1f, (boot || flat) ? 2.0f : lastfg ? 0.5f : 1.0f)), 0.1f);
        const double rr = (double)__fadd_rn(1.0f, __fmul_rn(dmin, 2.0f));
        if((double)R < __dmul_rn(rr, rr)) R = __fadd_rn(R, __fmul_rn(0.01f, __fsub_rn(V, 0.1f)));
        else R = fmaxf(__fsub_rn(R, __fdiv_rn(0.01f, V)), 1.0f);

        A.maps[pix * 2] = make_float4(T, R, V, 0.0f);
        A.maps[pix * 2 + 1] = make_float4(DminLT, DminST, rawLT, rawST);
        ((Col*)A.last_color)[pix] = cur_pack;
        ((Desc*)A.last_desc)[pix] = intra_pack;
    }

    const uint32_t b_raw = __ballot_sync(0xFFFFFFFFu, seg);
    const uint32_t b_unst = __ballot_sync(0xFFFFFFFFu, unstable_new);
    const uint32_t b_did = __ballot_sync(0xFFFFFFFFu, did);
    const uint32_t b_gop = __ballot_sync(0xFFFFFFFFu, has_gop);
    const uint32_t b_flat = __ballot_sync(0xFFFFFFFFu, flat);
    uint32_t b_int = 0;
#pragma unroll
    for(int d = 0; d < 5; ++d) {
        const uint32_t b = __ballot_sync(0xFFFFFFFFu, has_intent && intent_row == d);
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

/// illumination mask of the next frame: new[p] = did[p+1] ? (roi[p]==255) : did[p]  (snapshot semantics)
__global__ void __launch_bounds__(256) pawcs_illum_kernel(const PawArgs A) {
    const int wi = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if(wi >= A.WW) return;
    const size_t i = (size_t)y * A.WW + wi;
    const uint32_t d = A.did_bits[i], nxt = wi + 1 < A.WW ? A.did_bits[i + 1] : 0u;
    const uint32_t dn = (d >> 1) | (nxt << 31); // bit x = did[x+1]
    A.illum_bits[i] = ((dn & A.roi255_bits[i]) | (~dn & d)) & A.roi_bits[i];
}

/// global dictionary, step 1: the first pixel (raster order) that asked for it replaces the last word of the dictionary
template<int CH>
__global__ void __launch_bounds__(1024) pawcs_gword_replace(const PawArgs A) {
    GDict* gd = A.gd;
    const uint32_t win = gd->rep_winner;
    if(win == 0xFFFFFFFFu) return;
    const int g = gd->dict[A.NG - 1];
    float* m = A.gmap + (size_t)g * A.gW * A.gH;
    for(int i = threadIdx.x; i < A.gW * A.gH; i += blockDim.x) m[i] = 0.0f;
    if(threadIdx.x == 0) {
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
__global__ void __launch_bounds__(128) pawcs_gword_finish(const PawArgs A) {
    GDict* gd = A.gd;
    const int g = threadIdx.x;
    if(g < A.NG) {
        const unsigned long long a = gd->acc[g];
        if(a) { gd->weight[g] = (float)((double)gd->weight[g] + (double)(long long)a / 4294967296.0); gd->acc[g] = 0ull; }
    }
    if(g == 0) gd->rep_winner = 0xFFFFFFFFu;
}

/// Phase B: queued neighbour-dictionary updates (PAWCS.cpp:1164-1247), gathered per TARGET pixel, raster order of the source
#ifndef PAWB_MIN_BLOCKS
#define PAWB_MIN_BLOCKS 5   // 3 / 4 / 5 / 6 / 8 CTAs per SM -> 2.52 / 2.46 / 2.41 / 2.46 / 2.52 ms per 1080p frame
#endif
template<int CH>
__global__ void __launch_bounds__(256, PAWB_MIN_BLOCKS) pawcs_phaseB(const PawArgs A) {
    typedef typename Pack<CH>::Col Col;
    typedef typename Pack<CH>::Desc Desc;
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if(x < 2 || y < 2 || x > A.W - 3 || y > A.H - 3) return;
    const int wi = x >> 5, xb = x & 31;
    const FrameCtl* ctl = A.ctl; const GDict* gd = A.gd;
    const uint32_t frame = ctl->frame_idx, cooldown = ctl->cooldown, woff = gd->weight_offset;
    const bool boot = gd->boot != 0;
    const uint32_t flatK = CH == 1 ? 2u : 4u;
    const size_t pix = (size_t)y * A.Wp + x;
    const float init_w = __fdiv_rn(1.0f, (float)woff);
    // pass 1: which of the 25 possible sources aim at this pixel (bit = window position in raster order of the source)
    uint32_t hits = 0;
#pragma unroll
    for(int dy = -2; dy <= 2; ++dy) {
        const int qy = y + dy;
        const uint32_t* row = A.intent_bits + (size_t)(2 - dy) * A.bitplane + (size_t)qy * A.WW;
        const uint32_t left = wi > 0 ? row[wi - 1] : 0u, cur = row[wi], right = wi + 1 < A.WW ? row[wi + 1] : 0u;
        const unsigned long long lo = ((unsigned long long)cur << 32) | left, hi = ((unsigned long long)right << 32) | cur;
        uint32_t win = (xb >= 2) ? (uint32_t)(hi >> (xb - 2)) & 31u : (uint32_t)(lo >> (30 + xb)) & 31u;
        while(win) {
            const int k = __ffs(win) - 1;
            win &= win - 1;
            if((int)(A.intents[(size_t)qy * A.Wp + (x - 2 + k)].x & 0xFFu) == (2 - dy) * 5 + (4 - k)) hits |= 1u << ((dy + 2) * 5 + k);
        }
    }
    // pass 2: every lane pops ITS next hit, so the lanes of a warp walk their dictionaries together (as many rounds as the busiest
    // lane has hits) instead of one round per window position with a handful of lanes each
    {
        while(hits) {
            const int hi_ = __ffs(hits) - 1;
            hits &= hits - 1;
            const int r_ = hi_ / 5, k = hi_ - r_ * 5;
            const int qy = y + r_ - 2;
            const int qx = x - 2 + k;
            const size_t qpix = (size_t)qy * A.Wp + qx;
            const uint4 rec = A.intents[qpix];
            // source pixel (qx,qy) updates this pixel's dictionary with its own colour / descriptor / thresholds
            const uint32_t thrD = rec.x >> 8, thrC = rec.y, rate = rec.w;
            const float wthr = __uint_as_float(rec.z);
            const uchar* src = A.img + (size_t)qy * A.ipitch + (size_t)qx * CH;
            Col sc; uint32_t sc32;
            if constexpr (CH == 1) { sc = src[0]; sc32 = src[0]; } else { sc = (uint32_t)src[0] | ((uint32_t)src[1] << 8) | ((uint32_t)src[2] << 16); sc32 = sc; }
            const Desc sd = ((const Desc*)A.last_desc)[qpix];   // == the source's intra descriptor of this frame
            const Desc td = ((const Desc*)A.last_desc)[pix];    // target's intra descriptor of this frame
            const bool sflat = paw_bits(sd) < flatK;
            const uint32_t occ_incr = (1u + cooldown) << ((sflat || boot) ? 1 : 0);
            const bool traw = (A.raw_bits[y * A.WW + wi] >> xb) & 1u;
            const uint32_t src_id = (uint32_t)(qy * A.W + qx);
            float sum = 0.0f;
            // software-pipelined: the colour / descriptor of word j+1 are fetched while word j is tested (the loads are harmless
            // when the scan stops at j; a word rewritten by this very hit is never re-read by it)
            Col nbc = ((const Col*)A.lw_color)[pix];
            Desc nbd = ((const Desc*)A.lw_desc)[pix];
            for(int j = 0; j < A.NW && sum < wthr; ++j) {
                const size_t at = (size_t)j * A.plane + pix;
                const Col bc = nbc;
                const Desc bd = nbd;
                if(j + 1 < A.NW) { nbc = ((const Col*)A.lw_color)[at + A.plane]; nbd = ((const Desc*)A.lw_desc)[at + A.plane]; }
                uint32_t l1, cd;
                const uint32_t mix = paw_color_dist<CH>(sc32, col_as_u32(bc), l1, cd);
                const uint32_t hd = paw_hdist(sd, bd);
                const uint32_t incr = paw_bits(bd) < flatK ? occ_incr * 2u : occ_incr;
                bool credit = false, set_desc = false, set_col = false;
                if(mix <= thrC && hd <= thrD) credit = true;
                else if(!traw && sflat && (boot || (philox_draw(A.seed, frame, src_id, (uint32_t)j, DOM_PAWCS_B) % rate) == 0u)) {
                    const uint32_t lhd = paw_hdist(sd, td);
                    if(mix <= thrC && lhd <= thrD / 2u) { credit = true; set_desc = true; }
                    else if(CH != 1 && paw_bits(td) < flatK && lhd + hd <= thrD && cd <= thrC / 4u) { credit = true; set_col = true; }
                }
                if(credit) {
                    const uint32_t wf = A.lw_first[at], wl = A.lw_last[at], wo = A.lw_occ[at];
                    const float w = paw_weight(wf, wl, wo, frame, woff);
                    sum = __fadd_rn(sum, w);
                    if(CH != 1) { // Q8: the 1-channel path of the reference updates a by-value copy (PAWCS.cpp:838)
                        A.lw_last[at] = frame;
                        if(w < 1.0f) A.lw_occ[at] = wo + incr;
                        if(set_desc) ((Desc*)A.lw_desc)[at] = sd;
                        if(set_col) ((Col*)A.lw_color)[at] = sc;
                    }
                }
            }
            if(sum < init_w) {
                const size_t at = (size_t)(A.NW - 1) * A.plane + pix;
                ((Col*)A.lw_color)[at] = sc; ((Desc*)A.lw_desc)[at] = sd;
                A.lw_occ[at] = occ_incr; A.lw_first[at] = frame; A.lw_last[at] = frame;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// global maintenance (PAWCS.cpp:1300-1334): one CTA per global word, then the dictionary bubble pass, then the per-pixel LUTs
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) pawcs_gword_maintain(const PawArgs A, int recalc, int update) {
    __shared__ unsigned long long s_sum;
    __shared__ int s_zero, s_upd;
    GDict* gd = A.gd;
    const int g = blockIdx.x; // identity; every word is maintained exactly once whatever the dictionary order
    const int n = A.gW * A.gH;
    float* m = A.gmap + (size_t)g * n;
    float* tmp = A.gmap_tmp + (size_t)g * n;
    if(threadIdx.x == 0) { s_sum = 0ull; s_zero = 0; s_upd = 0; }
    __syncthreads();
    const bool live = gd->weight[g] > 0.0f;
    if(recalc && live) {
        long long acc = 0;
        for(int i = threadIdx.x; i < n; i += blockDim.x) acc += __double2ll_rn((double)m[i] * 4294967296.0);
#pragma unroll
        for(int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, o);
        if((threadIdx.x & 31) == 0) atomicAdd(&s_sum, (unsigned long long)acc);
        __syncthreads();
        if(threadIdx.x == 0) {
            float w = (float)((double)(long long)s_sum / 4294967296.0);
            if(w < 1.0f) { w = 0.0f; s_zero = 1; }
            gd->weight[g] = w;
            s_upd = w > 0.0f;
        }
        __syncthreads();
        if(s_zero) for(int i = threadIdx.x; i < n; i += blockDim.x) m[i] = 0.0f;
    } else {
        if(threadIdx.x == 0) s_upd = live;
        __syncthreads();
    }
    if(update && s_upd) {
        // accumulateProduct(map, -0.1, map, mask = nearest-downscaled ~dilate(lastFG)), weight *= 0.9, blur 3x3 replicate
        for(int i = threadIdx.x; i < n; i += blockDim.x) {
            const int cx = i % A.gW, cy = i / A.gW, x = cx * 2, y = cy * 2;
            if((A.dilinv_bits[y * A.WW + (x >> 5)] >> (x & 31)) & 1u) { const float v = m[i]; m[i] = __fadd_rn(v, __fmul_rn(v, -0.1f)); }
        }
        if(threadIdx.x == 0) gd->weight[g] = __fmul_rn(gd->weight[g], 0.9f);
        __syncthreads();
        for(int i = threadIdx.x; i < n; i += blockDim.x) {
            const int cx = i % A.gW, cy = i / A.gW;
            const int xl = max(cx - 1, 0), xr = min(cx + 1, A.gW - 1);
            double s = 0.0;
#pragma unroll
            for(int d = -1; d <= 1; ++d) {
                const float* r = m + (size_t)min(max(cy + d, 0), A.gH - 1) * A.gW;
                s += ((double)r[xl] + (double)r[cx]) + (double)r[xr];
            }
            tmp[i] = (float)(s * (1.0 / 9.0));
        }
        __syncthreads();
        for(int i = threadIdx.x; i < n; i += blockDim.x) m[i] = tmp[i];
    }
}
__global__ void pawcs_gdict_bubble(const PawArgs A) { // :1316-1317
    GDict* gd = A.gd;
    for(int i = 1; i < A.NG; ++i)
        if(gd->weight[gd->dict[i]] > gd->weight[gd->dict[i - 1]]) { const int t = gd->dict[i]; gd->dict[i] = gd->dict[i - 1]; gd->dict[i - 1] = t; }
}
/// one bubble pass over the per-pixel global-word LUT (:1319-1334 / :411-428); `only_if_refresh`: runs only when a refresh just happened
__global__ void __launch_bounds__(256) pawcs_glut_bubble(const PawArgs A, int only_if_refresh) {
    if(only_if_refresh && !A.gd->refresh_req) return;
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if(x >= A.W || y >= A.H) return;
    if(!((A.roi_bits[y * A.WW + (x >> 5)] >> (x & 31)) & 1u)) return;
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
    float tw = 0.0f, tc[CH], td[CH];
#pragma unroll
    for(int c = 0; c < CH; ++c) { tc[c] = 0.0f; td[c] = 0.0f; }
    for(int i = 0; i < A.NW; ++i) {
        const size_t at = (size_t)i * A.plane + pix;
        const float w = paw_weight(A.lw_first[at], A.lw_last[at], A.lw_occ[at], frame, woff);
        if(out_color) { const Col bc = ((const Col*)A.lw_color)[at];
#pragma unroll
            for(int c = 0; c < CH; ++c) tc[c] = __fadd_rn(tc[c], __fmul_rn((float)col_get(bc, c), w)); }
        if(out_desc) { const Desc bd = ((const Desc*)A.lw_desc)[at];
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
__global__ void __launch_bounds__(256) pawcs_tail1_kernel(const PawArgs A, int check_model) {
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
}
/// tail, part 2 (1 thread): reset logic (:1501-1515), next-frame factors
__global__ void pawcs_tail2_kernel(const PawArgs A) {
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
    auto valid = [&](int i) { const size_t at = (size_t)i * A.plane + pix; return !(A.lw_first[at] == 1u && A.lw_last[at] == 0u); };
    auto weight = [&](int i) { const size_t at = (size_t)i * A.plane + pix; return paw_weight(A.lw_first[at], A.lw_last[at], A.lw_occ[at], frame, woff); };
    if(decr > 0.0f)
        for(int i = 0; i < NW; ++i) { const size_t at = (size_t)i * A.plane + pix; if(valid(i)) { const uint32_t o = A.lw_occ[at]; A.lw_occ[at] = o - (uint32_t)__fmul_rn(decr, (float)o); } }
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
            if(paw_color_dist<CH>(col_as_u32(scol), col_as_u32(((const Col*)A.lw_color)[at]), l1, cd) <= thrC && paw_hdist(sdesc, ((const Desc*)A.lw_desc)[at]) <= thrD) {
                A.lw_occ[at] += 1u; A.lw_last[at] = frame; break;
            }
        }
        if(i == NW) {
            i = NW - 1;
            const size_t at = (size_t)i * A.plane + pix;
            ((Col*)A.lw_color)[at] = scol; ((Desc*)A.lw_desc)[at] = sdesc;
            A.lw_occ[at] = base_occ; A.lw_first[at] = frame; A.lw_last[at] = frame;
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
        const Col rc = ((const Col*)A.lw_color)[ar];
        if constexpr (CH == 1) ((Col*)A.lw_color)[at] = (uchar)min(max((int)rc + off, 0), 255);
        else ((Col*)A.lw_color)[at] = (uint32_t)min(max((int)(rc & 0xFFu) + off, 0), 255) | ((uint32_t)min(max((int)((rc >> 8) & 0xFFu) + off, 0), 255) << 8)
                                    | ((uint32_t)min(max((int)((rc >> 16) & 0xFFu) + off, 0), 255) << 16);
        ((Desc*)A.lw_desc)[at] = ((const Desc*)A.lw_desc)[ar];
        const uint32_t o = (uint32_t)__fmul_rn((float)A.lw_occ[ar], __fdiv_rn((float)(NW - i), (float)NW));
        A.lw_occ[at] = max(o, 1u); A.lw_first[at] = frame; A.lw_last[at] = frame;
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
                const Col bc = ((const Col*)A.lw_color)[pix]; const Desc bd = ((const Desc*)A.lw_desc)[pix];
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
                const float bw = paw_weight(A.lw_first[pix], A.lw_last[pix], A.lw_occ[pix], frame, woff);
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
/// runs after the conditional refresh kernels: bump the epoch they consumed
__global__ void pawcs_refresh_done(const PawArgs A) {
    if(A.gd->refresh_req) { A.ctl->refresh_epoch += 1; A.gd->refresh_req = PAW_REQ_NONE; A.gd->set_T_one = 0; }
}

} // namespace lvb
