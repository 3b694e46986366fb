// litiv_b200 — mask post-processing on bit-packed planes (1 bit / pixel), replacing the ~20 full-frame OpenCV
// calls of the reference (video/src/BackgroundSubtractorSuBSENSE.cpp:536-554, PAWCS.cpp:1443-1461,
// LOBSTER.cpp:578). Semantics per SURVEY.md Appendix E:
//   dilate/erode: rect kernel, pixels outside the image ignored; median on {0,255}: majority with replicated
//   border; floodFill((0,0)) + bitwise_not: background pixels not 4-connected to the image border ring.
#pragma once
#include "state.cuh"
#include <cooperative_groups.h>

namespace lvb {
namespace cg = cooperative_groups;

struct PostArgs {
    int W, H, WW, Wp;
    uint32_t* raw; uint32_t* lastraw; uint32_t* lastrawblink; uint32_t* blinks;
    uint32_t* tmpA; uint32_t* pre; uint32_t* reach; uint32_t* comb;
    uint32_t* lastfg; uint32_t* dilinv;
    uchar* out_mask; size_t out_pitch;
    float2* fin;
    FrameCtl* ctl;
    int median_k;
};

template<int FILL>
__device__ __forceinline__ void load3(const uint32_t* __restrict__ plane, int y, int wi, int H, int WW, int W, uint32_t& l, uint32_t& c, uint32_t& r, bool& valid) {
    valid = (y >= 0 && y < H);
    if(!valid) { l = c = r = (FILL == FILL_ONE) ? 0xFFFFFFFFu : 0u; return; }
    const uint32_t* row = plane + (size_t)y * WW;
    l = row_word<FILL>(row, wi - 1, WW, W); c = row_word<FILL>(row, wi, WW, W); r = row_word<FILL>(row, wi + 1, WW, W);
}
__device__ __forceinline__ uint32_t valid_mask(int wi, int WW, int W) {
    const int rem = W & 31;
    return (rem && wi == WW - 1) ? ((1u << rem) - 1u) : 0xFFFFFFFFu;
}

/// (2R+1)x(2R+1) rect dilate / erode of one mask word
template<int R, bool DILATE>
__device__ __forceinline__ uint32_t morph_word(const uint32_t* __restrict__ src, int y, int wi, int H, int WW, int W) {
    uint32_t acc = DILATE ? 0u : 0xFFFFFFFFu;
#pragma unroll
    for(int dy = -R; dy <= R; ++dy) {
        uint32_t l, c, r; bool v;
        load3<DILATE ? FILL_ZERO : FILL_ONE>(src, y + dy, wi, H, WW, W, l, c, r, v);
        const uint32_t h = hmorph<R, DILATE>(l, c, r);
        if(DILATE) acc |= h; else acc &= h;
    }
    return acc & valid_mask(wi, WW, W);
}

/// P1 (blink bookkeeping, SuBSENSE.cpp:536-539) + first half of MORPH_CLOSE (dilate 3x3, :540)
__global__ void __launch_bounds__(256) pp_blink_dilate(const PostArgs A) {
    const int wi = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if(wi >= A.WW) return;
    const size_t i = (size_t)y * A.WW + wi;
    const uint32_t raw = A.raw[i];
    const uint32_t cur_blink = raw ^ A.lastraw[i];
    A.blinks[i] = cur_blink | A.lastrawblink[i];
    A.lastrawblink[i] = cur_blink;
    A.lastraw[i] = raw;
    A.tmpA[i] = morph_word<1, true>(A.raw, y, wi, A.H, A.WW, A.W);
}
/// second half of MORPH_CLOSE (erode 3x3) + seeds of the border flood (the 2-px border ring is always background)
__global__ void __launch_bounds__(256) pp_erode_seed(const PostArgs A) {
    const int wi = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if(wi >= A.WW) return;
    const size_t i = (size_t)y * A.WW + wi;
    const uint32_t pre = morph_word<1, false>(A.tmpA, y, wi, A.H, A.WW, A.W);
    A.pre[i] = pre;
    const uint32_t vm = valid_mask(wi, A.WW, A.W);
    uint32_t seed = 0;
    if(y == 0 || y == A.H - 1) seed = vm;
    if(wi == 0) seed |= 1u;
    if(wi == A.WW - 1) seed |= 1u << ((A.W - 1) & 31);
    A.reach[i] = seed & ~pre & vm;
}

/// horizontal run fill of one 32-word chunk held one word per lane: returns the bits of m connected (towards higher
/// bit index, across lanes) to a seed bit. Carry-propagation trick: adding the seeds to the run mask ripples through
/// each run up to its end; the ripple across words is resolved with one ballot pair.
__device__ __forceinline__ uint32_t fill_up_chunk(uint32_t m, uint32_t s, uint32_t& carry) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t a = m + s;
    const bool g0 = a < m, p = (a == 0xFFFFFFFFu);
    const uint32_t G = __ballot_sync(0xFFFFFFFFu, g0), P = __ballot_sync(0xFFFFFFFFu, p);
    const uint32_t S = (G << 1) | carry;
    const unsigned long long sum = (unsigned long long)P + (unsigned long long)(S & P);
    const uint32_t C = S | ((uint32_t)sum ^ P);
    carry = (G >> 31) | (uint32_t)(sum >> 32);
    const uint32_t t = a + ((C >> lane) & 1u);
    return ((t ^ m) & m) | s;
}

constexpr int FLOOD_MAX_CHUNKS = 8; // rows up to 8192 pixels
constexpr int FLOOD_BAND = 16;

/// one row step: seeds = own reach | vertical neighbours' reach, restricted to background m; complete horizontal fill
__device__ __forceinline__ bool flood_row(const uint32_t* __restrict__ pre, uint32_t* reach, int y, int H, int WW, int W, int nchunks) {
    const uint32_t lane = threadIdx.x & 31;
    uint32_t m[FLOOD_MAX_CHUNKS], s[FLOOD_MAX_CHUNKS], old[FLOOD_MAX_CHUNKS], up[FLOOD_MAX_CHUNKS];
    bool changed = false;
#pragma unroll
    for(int k = 0; k < FLOOD_MAX_CHUNKS; ++k) {
        m[k] = 0; s[k] = 0; old[k] = 0;
        if(k < nchunks) {
            const int wi = k * 32 + lane;
            if(wi < WW) {
                const size_t i = (size_t)y * WW + wi;
                m[k] = ~pre[i] & valid_mask(wi, WW, W);
                old[k] = __ldcg(reach + i);
                uint32_t v = old[k];
                if(y > 0) v |= __ldcg(reach + i - WW);
                if(y < H - 1) v |= __ldcg(reach + i + WW);
                s[k] = v & m[k];
            }
        }
    }
    uint32_t carry = 0;
#pragma unroll
    for(int k = 0; k < FLOOD_MAX_CHUNKS; ++k) if(k < nchunks) up[k] = fill_up_chunk(m[k], s[k], carry);
    carry = 0;
#pragma unroll
    for(int rk = 0; rk < FLOOD_MAX_CHUNKS; ++rk) {
        const int k = nchunks - 1 - rk;
        if(rk < nchunks) {
            // reversed view: lane i holds the bit-reversed word of lane 31-i
            const uint32_t mr = __brev(__shfl_sync(0xFFFFFFFFu, m[k], 31 - lane));
            const uint32_t sr = __brev(__shfl_sync(0xFFFFFFFFu, s[k], 31 - lane));
            const uint32_t dr = fill_up_chunk(mr, sr, carry);
            const uint32_t down = __brev(__shfl_sync(0xFFFFFFFFu, dr, 31 - lane));
            const uint32_t res = up[k] | down;
            const int wi = k * 32 + lane;
            if(wi < WW && res != old[k]) { __stcg(reach + (size_t)y * WW + wi, res); changed = true; }
        }
    }
    return changed;
}

/// border flood to convergence (cv::floodFill((0,0)) equivalent). Cooperative launch; each warp owns a band of rows
/// and sweeps it down then up per global iteration; grid-wide barrier between iterations.
__global__ void __launch_bounds__(256) pp_flood(const PostArgs A, int bands) {
    cg::grid_group grid = cg::this_grid();
    const int warps_per_block = blockDim.x >> 5;
    const int gw = blockIdx.x * warps_per_block + (threadIdx.x >> 5), nw = gridDim.x * warps_per_block;
    const int nchunks = (A.WW + 31) >> 5;
    volatile uint32_t* flags = A.ctl->flood_changed; // 3 rotating flags (see below), all zero on entry
    __shared__ uint32_t s_flags[3];
    for(uint32_t it = 0;; ++it) {
        bool changed = false;
        for(int b = gw; b < bands; b += nw) {
            const int y0 = b * FLOOD_BAND, y1 = min(y0 + FLOOD_BAND, A.H);
            for(int y = y0; y < y1; ++y) changed |= flood_row(A.pre, A.reach, y, A.H, A.WW, A.W, nchunks);
            for(int y = y1 - 2; y >= y0; --y) changed |= flood_row(A.pre, A.reach, y, A.H, A.WW, A.W, nchunks);
        }
        if(__any_sync(0xFFFFFFFFu, changed) && (threadIdx.x & 31) == 0) atomicOr((uint32_t*)&flags[it % 3], 1u);
        __threadfence();
        grid.sync();
        if(threadIdx.x == 0) {
            s_flags[0] = flags[it % 3];
            // flag (it+2)%3 was last read after barrier it-1 and is next written after barrier it+1: safe to clear now
            if(blockIdx.x == 0) flags[(it + 2) % 3] = 0;
        }
        __syncthreads();
        const uint32_t f = s_flags[0];
        __syncthreads();
        if(!f) break;
    }
    // leave all three flags zero for the next frame: flag it%3 is 0 (loop exit), (it+2)%3 cleared above, (it+1)%3 untouched since cleared
}

/// m = raw | holes | erode7x7(pre)   (SuBSENSE.cpp:543-546); holes = background of `pre` not reached by the border flood
__global__ void __launch_bounds__(256) pp_combine(const PostArgs A) {
    const int wi = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if(wi >= A.WW) return;
    const size_t i = (size_t)y * A.WW + wi;
    const uint32_t vm = valid_mask(wi, A.WW, A.W);
    const uint32_t holes = ~A.pre[i] & ~A.reach[i] & vm;
    A.comb[i] = A.raw[i] | holes | morph_word<3, false>(A.pre, y, wi, A.H, A.WW, A.W);
}

/// k x k majority (== cv::medianBlur on a binary mask, replicated borders); one thread per pixel.
/// Writes the bit-packed result and the byte mask handed back to the caller.
__global__ void __launch_bounds__(256) pp_median(const uint32_t* __restrict__ src, uint32_t* dst, uchar* out_mask, size_t out_pitch,
                                                  int W, int H, int WW, int k) {
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    const int wi = x >> 5, xb = x & 31, r = k >> 1;
    bool on = false;
    if(x < W && y < H) {
        const uint32_t wmask = (1u << k) - 1u;
        int cnt = 0;
        for(int dy = -r; dy <= r; ++dy) {
            const int yy = clampi(y + dy, 0, H - 1);
            const uint32_t* row = src + (size_t)yy * WW;
            const uint32_t l = row_word<FILL_REPL>(row, wi - 1, WW, W), c = row_word<FILL_REPL>(row, wi, WW, W), rr = row_word<FILL_REPL>(row, wi + 1, WW, W);
            const unsigned long long lo = ((unsigned long long)c << 32) | l, hi = ((unsigned long long)rr << 32) | c;
            const uint32_t win = (xb >= r) ? (uint32_t)(hi >> (xb - r)) : (uint32_t)(lo >> (32 + xb - r));
            cnt += __popc(win & wmask);
        }
        on = cnt > (k * k) / 2;
        if(out_mask) out_mask[(size_t)y * out_pitch + x] = on ? 255 : 0;
    }
    const uint32_t b = __ballot_sync(0xFFFFFFFFu, on);
    if(threadIdx.x == 0 && y < H && wi < WW) dst[(size_t)y * WW + wi] = b;
}

/// dilate 7x7 of the final mask + blink gating (SuBSENSE.cpp:548-551)
__global__ void __launch_bounds__(256) pp_dilate_blink(const PostArgs A) {
    const int wi = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if(wi >= A.WW) return;
    const size_t i = (size_t)y * A.WW + wi;
    const uint32_t dil = morph_word<3, true>(A.lastfg, y, wi, A.H, A.WW, A.W);
    const uint32_t ninv = ~dil & valid_mask(wi, A.WW, A.W);
    A.blinks[i] = A.blinks[i] & A.dilinv[i] & ninv;
    A.dilinv[i] = ninv;
}

/// final-segmentation EMAs (SuBSENSE.cpp:553-554): cv::addWeighted accumulates in double and rounds once
__global__ void __launch_bounds__(256) pp_final_ema(const PostArgs A) {
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if(x >= A.W || y >= A.H) return;
    const float aLT = A.ctl->aLT, aST = A.ctl->aST;
    const double v = ((A.lastfg[(size_t)y * A.WW + (x >> 5)] >> (x & 31)) & 1u) ? 255.0 : 0.0;
    const size_t pix = (size_t)y * A.Wp + x;
    float2 f = A.fin[pix];
    f.x = (float)__dadd_rn(__dmul_rn((double)f.x, (double)__fsub_rn(1.0f, aLT)), __dmul_rn(v, __dmul_rn(1.0 / 255, (double)aLT)));
    f.y = (float)__dadd_rn(__dmul_rn((double)f.y, (double)__fsub_rn(1.0f, aST)), __dmul_rn(v, __dmul_rn(1.0 / 255, (double)aST)));
    A.fin[pix] = f;
}

} // namespace lvb
