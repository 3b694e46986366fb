// litiv_b200 — LOBSTER per-frame kernel (replaces BackgroundSubtractorLOBSTER_<NonParallel>::apply,
// reference video/src/BackgroundSubtractorLOBSTER.cpp:459-581), the dense LBSP extractor
// (LBSP::compute2, features2d/src/LBSP.cpp:102-152) and the shared initialisation / background-image kernels.
#pragma once
#include "subsense.cuh"
#include "edge_px.cuh"

namespace lvb {

#ifndef LOB_MIN_BLOCKS
#define LOB_MIN_BLOCKS 1
#endif
// sample records in flight per pixel (sweep: gray 2 / 4 / 8 -> 105.7 / 101.0 / 104.3 us per 1080p frame, RGB 119.5 / 127.1 / 175.1 us)
#ifndef LOB_CHUNK
#define LOB_CHUNK (CH == 1 ? 4 : 2)
#endif
/// T7: every LBSP threshold of the LUT is <= 127 (true for the reference's defaults): the 7-bit compare of lbsp_threshold
template<int CH, bool T7>
__global__ void __launch_bounds__(TILE_W * TILE_H, LOB_MIN_BLOCKS)
lobster_phaseA(const SubArgs A, const __grid_constant__ CUtensorMap tmap) {
    typedef typename Pack<CH>::Col Col;
    typedef typename Pack<CH>::Desc Desc;
    typedef typename Pack<CH>::Rec Rec;
    constexpr int PITCH = tile_pitch(CH);
    __shared__ __align__(128) uchar s_tile[PITCH * TILE_ROWS];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ uchar s_lut[256];
    __shared__ uint32_t s_cnt[4];

    const int x0 = blockIdx.x * TILE_W, y0 = blockIdx.y * TILE_H;
    const int tid = threadIdx.y * TILE_W + threadIdx.x;
    for(int i = tid; i < 256; i += TILE_W * TILE_H) s_lut[i] = A.lut[i];
    if(tid < 4) s_cnt[tid] = 0;
    stage_tile<CH>(s_tile, &s_bar, &tmap, A.use_tma, A.img, A.ipitch, A.W, A.H, x0, y0);
    __syncthreads();

    const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
    const bool in_img = (x < A.W) && (y < A.H);
    const int wi = y * A.WW + (x >> 5);
    const uint32_t lane_bit = 1u << (x & 31);
    const uint32_t w_roi = (y < A.H && (x >> 5) < A.WW) ? A.roi_bits[wi] : 0u;
    const bool active = in_img && (w_roi & lane_bit);
    const size_t pix = (size_t)y * A.Wp + x;
    const int sx = threadIdx.x + HALO, sy = threadIdx.y + HALO;
    bool is_fg = false, has_intent = false;
    uint32_t scanned = 0, writes = 0;
    uint32_t intent = NO_INTENT;

    uint32_t cur[CH];
    Col cur_pack;
#pragma unroll
    for(int c = 0; c < CH; ++c) cur[c] = s_tile[tile_shift(CH) + sy * PITCH + sx * CH + c];
    if constexpr (CH == 1) cur_pack = (uchar)cur[0]; else cur_pack = cur[0] | (cur[1] << 8) | (cur[2] << 16);

    if(active) {
        const uint32_t N = (uint32_t)A.N, REQ = (uint32_t)A.REQ;
        const uint32_t colorThr = (uint32_t)A.min_color, descThr = (uint32_t)A.desc_off;
        const uint32_t totD = descThr * 3u, totC = colorThr * 3u, scD = totD >> 1, scC = totC >> 1;
        Lookup16 L[CH];
#pragma unroll
        for(int c = 0; c < CH; ++c) L[c] = lbsp_lookup_smem<CH>(s_tile + tile_shift(CH), PITCH, sx, sy, c);
        const Rec* bgr = (const Rec*)A.bg + pix;
        uint32_t good = 0, s = 0;
        // LOBSTER.cpp:481-495 / :533-553: "while(good < REQ && s < N)". The records of CHUNK samples are in flight together (a foreground
        // pixel scans all N: one dependent DRAM round trip per sample held whole CTAs for N round trips); the tests run in sample order and
        // stop exactly where the reference's loop stops
        constexpr int CHUNK = LOB_CHUNK;
        for(uint32_t s0 = 0; s0 < N && good < REQ; s0 += CHUNK) {
            Rec recs[CHUNK];
#pragma unroll
            for(int k = 0; k < CHUNK; ++k) recs[k] = s0 + k < N ? bgr[(size_t)(s0 + k) * A.plane] : Rec();
#pragma unroll
            for(int k = 0; k < CHUNK; ++k) {
                if(s0 + k < N && good < REQ) {
                    const Rec rec = recs[k];
                    const Col bc = rec_col(rec);
                    bool ok = true;
                    uint32_t tc = 0;
#pragma unroll
                    for(int c = 0; c < CH; ++c) {
                        const uint32_t b = col_get(bc, c);
                        const uint32_t cd = cur[c] > b ? cur[c] - b : b - cur[c];
                        ok = ok && (cd <= (CH == 1 ? colorThr / 2u : scC));
                        tc += cd;
                    }
                    if(ok) {
                        const Desc bd = rec_desc(rec);
                        uint32_t td = 0;
#pragma unroll
                        for(int c = 0; c < CH; ++c) {
                            const uint32_t b = col_get(bc, c);
                            const uint32_t dd = __popc(lbsp_threshold<T7>(L[c], b, s_lut[b]) ^ desc_get(bd, c));
                            ok = ok && (dd <= (CH == 1 ? descThr : scD));
                            td += dd;
                        }
                        if(CH != 1) ok = ok && (td <= totD) && (tc <= totC);
                        if(ok) ++good;
                    }
                    ++s;
                }
            }
        }
        scanned = s;
        if(good < REQ) is_fg = true;
        else {
            const uint32_t frame = A.ctl->frame_idx;
            const uint32_t pixid = (uint32_t)(y * A.W + x);
            const uint32_t LR = A.lr_fixed;
            const uint4 rnd = philox_block(A.seed, frame, pixid, 0, DOM_APPLY);
            const bool own = (rnd.x % LR) == 0, nb = (rnd.z % LR) == 0;
            if(own || nb) {
                uint32_t intra[CH];
#pragma unroll
                for(int c = 0; c < CH; ++c) intra[c] = lbsp_threshold<T7>(L[c], cur[c], s_lut[cur[c]]);
                Desc intra_pack;
                if constexpr (CH == 1) intra_pack = (ushort)intra[0]; else intra_pack = make_uint2(intra[0] | (intra[1] << 16), intra[2]);
                if(own) {
                    const uint32_t slot = rnd.y % N;
                    ((Rec*)A.bg)[(size_t)slot * A.plane + pix] = rec_make(cur_pack, intra_pack);
                    ++writes;
                }
                if(nb) {
                    int dx, dy;
                    neighbor_offset(true, rnd.w, dx, dy);
                    const int nx = clampi(x + dx, 2, A.W - 3), ny = clampi(y + dy, 2, A.H - 3);
                    const uint32_t slot = (rnd.y / N) % N; // one Philox block per pixel
                    intent = (((ny - y + 2) * 5 + (nx - x + 2)) << 8) | slot;
                    ((Desc*)A.last_desc)[pix] = intra_pack; // scratch plane read by phase B (not the reference's m_oLastDescFrame)
                    has_intent = true;
                }
            }
        }
    }
    if(in_img) ((Col*)A.last_color)[pix] = cur_pack; // whole frame (:580)

    const uint32_t b_raw = __ballot_sync(0xFFFFFFFFu, is_fg);
    if(in_img) A.intents[pix] = (ushort)intent; // every pixel, every frame (phase B scans the plane without a has-intent mask)
    if(y < A.H && (x >> 5) < A.WW) {
        if(threadIdx.x == 0) A.raw_bits[wi] = b_raw;
    }
    if(A.collect_stats) {
        uint32_t sc = scanned, wr = writes + (has_intent ? 1u : 0u);
#pragma unroll
        for(int o = 16; o > 0; o >>= 1) { sc += __shfl_xor_sync(0xFFFFFFFFu, sc, o); wr += __shfl_xor_sync(0xFFFFFFFFu, wr, o); }
        if(threadIdx.x == 0) { atomicAdd(&s_cnt[1], sc); atomicAdd(&s_cnt[2], wr); atomicAdd(&s_cnt[3], __popc(b_raw)); }
        __syncthreads();
        if(tid == 0) {
            atomicAdd(&A.ctl->stat_scanned, (unsigned long long)s_cnt[1]);
            atomicAdd(&A.ctl->stat_writes, (unsigned long long)s_cnt[2]);
            atomicAdd(&A.ctl->stat_fg, (unsigned long long)s_cnt[3]);
        }
    }
}

/// initialisation (BackgroundSubtractionUtils.cpp:117-154 + BackgroundSubtractorLBSP.cpp:21-65):
/// last_color = init image inside the ROI, last_desc = intra LBSP for ROI pixels strictly inside the 2-px border + 1 (Q4)
struct InitArgs {
    int W, H, Wp, WW;
    const uchar* img; size_t ipitch;
    void* last_color; void* last_desc;
    const uint32_t* roi_bits; const uchar* lut;
    int use_tma;
};
template<int CH>
__global__ void __launch_bounds__(TILE_W * TILE_H) init_frame_kernel(const InitArgs A, const __grid_constant__ CUtensorMap tmap) {
    typedef typename Pack<CH>::Col Col;
    typedef typename Pack<CH>::Desc Desc;
    typedef typename Pack<CH>::Rec Rec;
    constexpr int PITCH = tile_pitch(CH);
    __shared__ __align__(128) uchar s_tile[PITCH * TILE_ROWS];
    __shared__ __align__(8) uint64_t s_bar;
    const int x0 = blockIdx.x * TILE_W, y0 = blockIdx.y * TILE_H;
    stage_tile<CH>(s_tile, &s_bar, &tmap, A.use_tma, A.img, A.ipitch, A.W, A.H, x0, y0);
    __syncthreads();
    const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
    if(x >= A.W || y >= A.H) return;
    const bool roi = (A.roi_bits[y * A.WW + (x >> 5)] >> (x & 31)) & 1u;
    const int sx = threadIdx.x + HALO, sy = threadIdx.y + HALO;
    uint32_t cur[CH], d[CH];
#pragma unroll
    for(int c = 0; c < CH; ++c) {
        cur[c] = roi ? s_tile[tile_shift(CH) + sy * PITCH + sx * CH + c] : 0u;
        d[c] = 0;
        if(roi && x > 2 && y > 2 && x < A.W - 2 && y < A.H - 2) {
            const Lookup16 L = lbsp_lookup_smem<CH>(s_tile + tile_shift(CH), PITCH, sx, sy, c);
            d[c] = lbsp_threshold(L, cur[c], A.lut[cur[c]]);
        }
    }
    const size_t pix = (size_t)y * A.Wp + x;
    if constexpr (CH == 1) { ((Col*)A.last_color)[pix] = (uchar)cur[0]; ((Desc*)A.last_desc)[pix] = (ushort)d[0]; }
    else { ((Col*)A.last_color)[pix] = cur[0] | (cur[1] << 8) | (cur[2] << 16); ((Desc*)A.last_desc)[pix] = make_uint2(d[0] | (d[1] << 16), d[2]); }
}

/// getBackgroundImage / getBackgroundDescriptorsImage (SuBSENSE.cpp:614-649, LOBSTER.cpp:583-620):
/// float mean accumulated sample by sample (x/N each), converted round-half-even with saturation
template<int CH>
__global__ void __launch_bounds__(256) background_image_kernel(const void* bg, size_t plane, int N, int W, int H, int Wp,
                                                                uchar* out_color, ushort* out_desc) {
    typedef typename Pack<CH>::Col Col;
    typedef typename Pack<CH>::Desc Desc;
    typedef typename Pack<CH>::Rec Rec;
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if(x >= W || y >= H) return;
    const size_t pix = (size_t)y * Wp + x, o = ((size_t)y * W + x) * CH;
    float acc[CH];
#pragma unroll
    for(int c = 0; c < CH; ++c) acc[c] = 0.f;
    for(int s = 0; s < N; ++s) {
        if(out_color) {
            const Col v = rec_col(((const Rec*)bg)[(size_t)s * plane + pix]);
#pragma unroll
            for(int c = 0; c < CH; ++c) acc[c] = __fadd_rn(acc[c], __fdiv_rn((float)col_get(v, c), (float)N));
        } else {
            const Desc v = rec_desc(((const Rec*)bg)[(size_t)s * plane + pix]);
#pragma unroll
            for(int c = 0; c < CH; ++c) acc[c] = __fadd_rn(acc[c], __fdiv_rn((float)desc_get(v, c), (float)N));
        }
    }
#pragma unroll
    for(int c = 0; c < CH; ++c) {
        if(out_color) out_color[o + c] = (uchar)fminf(fmaxf(rintf(acc[c]), 0.f), 255.f);
        else out_desc[o + c] = (ushort)fminf(fmaxf(rintf(acc[c]), 0.f), 65535.f);
    }
}

/// dense LBSP (LBSP::compute2): intra (ref == image) or inter (separate reference image), absolute or relative threshold
struct LbspArgs {
    int W, H;
    const uchar* img; size_t ipitch;
    const uchar* ref; size_t rpitch;   // null -> intra
    ushort* out;                       // [H][W][CH], the 2-px border is left untouched
    int use_rel; float rel; int thr;
    int use_tma;
};
template<int CH>
__global__ void __launch_bounds__(TILE_W * TILE_H) lbsp_dense_kernel(const LbspArgs A, const __grid_constant__ CUtensorMap tmap) {
    constexpr int PITCH = tile_pitch(CH);
    __shared__ __align__(128) uchar s_tile[PITCH * TILE_ROWS];
    __shared__ __align__(8) uint64_t s_bar;
    const int x0 = blockIdx.x * TILE_W, y0 = blockIdx.y * TILE_H;
    stage_tile<CH>(s_tile, &s_bar, &tmap, A.use_tma, A.img, A.ipitch, A.W, A.H, x0, y0);
    __syncthreads();
    const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
    if(x < 2 || y < 2 || x >= A.W - 2 || y >= A.H - 2) return;
    const int sx = threadIdx.x + HALO, sy = threadIdx.y + HALO;
    const uint32_t tabs = (uint32_t)min(max(A.thr, 0), 255);
#pragma unroll
    for(int c = 0; c < CH; ++c) {
        const uint32_t ref = A.ref ? A.ref[(size_t)y * A.rpitch + x * CH + c] : s_tile[tile_shift(CH) + sy * PITCH + sx * CH + c];
        uint32_t t = tabs;
        if(A.use_rel) t = (uint32_t)fminf(fmaxf(rintf(__fadd_rn(__fmul_rn((float)ref, A.rel), (float)A.thr)), 0.f), 255.f);
        const Lookup16 L = lbsp_lookup_smem<CH>(s_tile + tile_shift(CH), PITCH, sx, sy, c);
        A.out[((size_t)y * A.W + x) * CH + c] = (ushort)lbsp_threshold(L, ref, t);
    }
}

/// dense LBSP gradient map (LBSP::computeDescriptor_gradient<C, 20, 2>, features2d/include/litiv/features2d/LBSP.hpp:235-256, the
/// per-pixel primitive of imgproc/src/EdgeDetectorLBSP.cpp:253): 4 bytes per pixel = gradX (int8), gradY (int8), magnitude, 0.
/// Pixels within the 2-px border get (0,0,0,0), the value the edge detector's all-equal border lookup yields (:84-100).
/// `combine` (EdgeDetectorLBSP, edge_px.cuh): the value written is edge_combine(gradient, coarser level's map at (y/2, x/2)), or with the
/// detector's initial value when `coarse` is null (the coarsest level)
struct LbspGradArgs { int W, H; const uchar* img; size_t ipitch; uchar4* out; int use_tma; int combine; const uchar4* coarse; int Wc; };
template<int CH>
__global__ void __launch_bounds__(TILE_W * TILE_H) lbsp_gradient_kernel(const LbspGradArgs A, const __grid_constant__ CUtensorMap tmap) {
    constexpr int PITCH = tile_pitch(CH);
    __shared__ __align__(128) uchar s_tile[PITCH * TILE_ROWS];
    __shared__ __align__(8) uint64_t s_bar;
    const int x0 = blockIdx.x * TILE_W, y0 = blockIdx.y * TILE_H;
    stage_tile<CH>(s_tile, &s_bar, &tmap, A.use_tma, A.img, A.ipitch, A.W, A.H, x0, y0);
    __syncthreads();
    const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
    if(x >= A.W || y >= A.H) return;
    uchar4 o = make_uchar4(0, 0, 0, 0);
    if(x >= 2 && y >= 2 && x < A.W - 2 && y < A.H - 2) {
        const int sx = threadIdx.x + HALO, sy = threadIdx.y + HALO;
        uint32_t best = 0; int best_mag = -1;
#pragma unroll
        for(int k = 0; k < CH; ++k) {
            const int c = k == 0 ? CH - 1 : k - 1; // the reference starts from the last channel and replaces on strictly greater
            const uint32_t ref = s_tile[tile_shift(CH) + sy * PITCH + sx * CH + c];
            const Lookup16 L = lbsp_lookup_smem<CH>(s_tile + tile_shift(CH), PITCH, sx, sy, c);
            const uint32_t d = lbsp_threshold<true>(L, ref, ((ref >> 2) + 20u) >> 1);   // threshold <= 41: the 7-bit compare path
            const int m = __popc(d);
            if(best_mag < m) { best_mag = m; best = d; }
        }
        constexpr uint32_t XP = (1u<<0)|(1u<<4)|(1u<<7)|(1u<<9)|(1u<<12)|(1u<<15), XN = (1u<<1)|(1u<<5)|(1u<<6)|(1u<<11)|(1u<<13)|(1u<<14);
        constexpr uint32_t YP = (1u<<3)|(1u<<4)|(1u<<6)|(1u<<8)|(1u<<13)|(1u<<15), YN = (1u<<2)|(1u<<5)|(1u<<7)|(1u<<10)|(1u<<12)|(1u<<14); // LBSP.hpp:288-291
        o.x = (uchar)(signed char)(__popc(best & XP) - __popc(best & XN));
        o.y = (uchar)(signed char)(__popc(best & YP) - __popc(best & YN));
        o.z = (uchar)best_mag;
    }
    if(A.combine) o = lvb_edge::edge_combine(o, A.coarse ? A.coarse[(size_t)(y >> 1) * A.Wc + (x >> 1)] : lvb_edge::edge_init_value());
    A.out[(size_t)y * A.W + x] = o;
}

} // namespace lvb
