// litiv_b200 — kernels of the LBSP edge detector (reference imgproc/src/EdgeDetectorLBSP.cpp:47-433). The per-pixel
// bodies live in edge_px.cuh (shared with the CPU emulation of the tests); the per-level LBSP gradient is lbsp_gradient_kernel of
// lobster.cuh (TMA-staged tile + 5x5 halo). All maps are HBM-resident per detector object. Algorithmic bytes of one apply_threshold
// call per full-resolution pixel, three levels (sizes 1 + 1/4 + 1/16 = 1.31): image read C and 4-byte map written, re-read and rewritten by
// the combination (C + 13 per level pixel, + C for writing the two coarser images), the suppression's map read 4 and mask write 1 (its 5x5
// window is served by L1 / L2), one mask read per flood sweep, mask read + result write 2: about 33 B/px for RGB with four sweeps
// (tools/bench_edge.py). The detector is launch bound at CDnet sizes (10 launches per call, no host round trip).
#pragma once
#include "edge_px.cuh"

namespace lvb_edge {

/// one thread per (pixel, channel) of the next pyramid level
__global__ void __launch_bounds__(256) edge_pyr_down_kernel(const uchar* cur, size_t cpitch, int Wc, int Hc, int C, uchar* nxt, size_t npitch, int Wn, int Hn) {
    const int xk = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if(xk >= Wn * C || y >= Hn) return;
    const int x = xk / C, k = xk - x * C;
    nxt[(size_t)y * npitch + xk] = pyr_down_px(cur, cpitch, Wc, Hc, C, 2 * y, 2 * x, k);
}

// (V_l = combine(own gradient, coarser V or the initial value) is the epilogue of lbsp_gradient_kernel, lobster.cuh)

/// mask rows 0 .. H-3 <- suppression class of gradient rows 2 .. H-1; rows H-2, H-1 are left as the previous call left them.
/// E / S (optional): the bit planes of the hysteresis (E: maybe or edge, S: edge), all H rows, one word per warp (blockDim = (32, 8))
__global__ void __launch_bounds__(256) edge_nms_kernel(const EdgeMaps m, unsigned lo, unsigned hi, uchar* mask, uint32_t* E, uint32_t* S, int WW) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if(y >= m.H) return;
    uchar v = EDGE_NONE;
    if(x < m.W) {
        if(y < m.H - 2) { v = edge_mask_value(m, y, x, lo, hi, mask); mask[(size_t)y * m.W + x] = v; }
        else v = mask[(size_t)y * m.W + x];
    }
    if(E) {
        const uint32_t e = __ballot_sync(0xFFFFFFFFu, v == EDGE_MAYBE || v == EDGE_YES), sd = __ballot_sync(0xFFFFFFFFu, v == EDGE_YES);
        if(threadIdx.x == 0) { E[(size_t)y * WW + blockIdx.x] = e; S[(size_t)y * WW + blockIdx.x] = sd; }
    }
}

/// hysteresis (:353-372), round-1 form kept for cross-checks (LVB_EDGE_SWEEPS=1): one relaxation sweep. Every CTA iterates its 32x16 tile (+1 halo) in shared memory until the tile is stable
/// (bounded), then writes back the pixels it turned from "maybe" to "edge" and raises *changed. The host repeats sweeps until a sweep
/// changes nothing; transitions are monotone (0 -> 2 only), so concurrent tiles reading each other's halo mid-sweep is harmless.
constexpr int FL_W = 32, FL_H = 16, FL_LOCAL_ITERS = 64;
__global__ void __launch_bounds__(FL_W * FL_H) edge_flood_kernel(uchar* mask, int W, int H, int* changed) {
    __shared__ uchar s[FL_H + 2][FL_W + 2];
    const int x0 = blockIdx.x * FL_W, y0 = blockIdx.y * FL_H, tx = threadIdx.x, ty = threadIdx.y;
    for(int i = ty * FL_W + tx; i < (FL_H + 2) * (FL_W + 2); i += FL_W * FL_H) {
        const int sy = i / (FL_W + 2), sx = i - sy * (FL_W + 2), gx = x0 + sx - 1, gy = y0 + sy - 1;
        s[sy][sx] = (gx >= 0 && gy >= 0 && gx < W && gy < H) ? mask[(size_t)gy * W + gx] : (uchar)EDGE_NONE;
    }
    __syncthreads();
    const int x = x0 + tx, y = y0 + ty;
    const bool inside = x < W && y < H;
    bool mine = inside && s[ty + 1][tx + 1] == EDGE_MAYBE, turned = false;
    for(int it = 0; it < FL_LOCAL_ITERS; ++it) {
        bool set = false;
        if(mine) {
            set = s[ty][tx] == EDGE_YES || s[ty][tx + 1] == EDGE_YES || s[ty][tx + 2] == EDGE_YES || s[ty + 1][tx] == EDGE_YES ||
                  s[ty + 1][tx + 2] == EDGE_YES || s[ty + 2][tx] == EDGE_YES || s[ty + 2][tx + 1] == EDGE_YES || s[ty + 2][tx + 2] == EDGE_YES;
        }
        __syncthreads();
        if(set) { s[ty + 1][tx + 1] = EDGE_YES; mine = false; turned = true; }
        if(!__syncthreads_or(set ? 1 : 0)) break;
    }
    if(turned) mask[(size_t)y * W + x] = EDGE_YES;
    if(__syncthreads_or(turned ? 1 : 0) && tx == 0 && ty == 0) *changed = 1;
}

// ---- hysteresis as connected components (replaces the relaxation sweeps above on the product path) ---------------------------------------
// The flood turns every "maybe" pixel that is 8-connected, through "maybe" / "edge" pixels, to an "edge" pixel into an edge. That is: the
// connected components (8-neighbourhood) of E = {maybe, edge} that contain a seed S = {edge}. Same machinery as the hole filling of the
// SuBSENSE mask chain (postproc.cuh): nodes = ROW RUNS of E, lock-free union-find, node 0 = "holds a seed"; three contact masks per row pair
// (straight, two diagonals) each give one union per maximal run of the contact. Cost independent of the edge geometry, no host round trip
// (the sweeps needed 12 launches and three flag read-backs at 1080p).
struct EdgeUF { int W, H, WW, RS; uint32_t* E; uint32_t* S; uint32_t* parent; ushort* rankbase; uchar* mask; };

// (the bit planes E / S are written by edge_nms_kernel)
__device__ __forceinline__ uint32_t euf_word(const uint32_t* __restrict__ P, int y, int wi, int WW) { return (wi < 0 || wi >= WW) ? 0u : P[(size_t)y * WW + wi]; }
/// node id of the E run of row y that contains bit b of word wi
__device__ __forceinline__ uint32_t euf_run_id(const EdgeUF& A, int y, int wi, int b) {
    const uint32_t st = lvb::run_starts(euf_word(A.E, y, wi, A.WW), euf_word(A.E, y, wi - 1, A.WW));
    const uint32_t upto = b == 31 ? 0xFFFFFFFFu : ((2u << b) - 1u);
    return 1u + (uint32_t)y * A.RS + A.rankbase[(size_t)y * A.WW + wi] + __popc(st & upto) - 1u;
}
/// per row (one warp): number the runs of E, every run its own root
__global__ void __launch_bounds__(256) edge_uf_init(const EdgeUF A) {
    const int y = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if(y >= A.H) return;
    const int nchunks = (A.WW + 31) >> 5;
    uint32_t base = 0;
    for(int k = 0; k < nchunks; ++k) {
        const int wi = k * 32 + lane;
        const uint32_t st = lvb::run_starts(euf_word(A.E, y, wi, A.WW), euf_word(A.E, y, wi - 1, A.WW));
        uint32_t cnt = __popc(st), incl = cnt;
#pragma unroll
        for(int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, o); if(lane >= o) incl += v; }
        const uint32_t excl = base + incl - cnt;
        if(wi < A.WW) {
            A.rankbase[(size_t)y * A.WW + wi] = (ushort)excl;
            for(uint32_t r = 0; r < cnt; ++r) { const uint32_t id = 1u + (uint32_t)y * A.RS + excl + r; A.parent[id] = id; }
        }
        base += __shfl_sync(0xFFFFFFFFu, incl, 31);
    }
    if(y == 0 && threadIdx.x == 0) A.parent[0] = 0u;
}
/// per row (one warp): runs that hold a seed go under node 0; runs of rows y-1 and y that touch (8-neighbourhood) are united
__global__ void __launch_bounds__(256) edge_uf_union(const EdgeUF A) {
    const int y = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if(y >= A.H) return;
    for(int wi = lane; wi < A.WW; wi += 32) {
        const uint32_t e = euf_word(A.E, y, wi, A.WW), ep = euf_word(A.E, y, wi - 1, A.WW);
        {   // seeds: one union per maximal run of S (S is a subset of E)
            const uint32_t sd = euf_word(A.S, y, wi, A.WW);
            uint32_t ss = lvb::run_starts(sd, euf_word(A.S, y, wi - 1, A.WW));
            while(ss) { const int b = __ffs(ss) - 1; ss &= ss - 1; lvb::uf_union(A.parent, euf_run_id(A, y, wi, b), 0u); }
        }
        if(y == 0) continue;
        const uint32_t u = euf_word(A.E, y - 1, wi, A.WW), ul = euf_word(A.E, y - 1, wi - 1, A.WW), ur = euf_word(A.E, y - 1, wi + 1, A.WW);
        const uint32_t ull = euf_word(A.E, y - 1, wi - 2, A.WW);
        // contact masks in row-y coordinates: bit x set when pixel (y,x) touches pixel (y-1, x+d)
#pragma unroll
        for(int d = -1; d <= 1; ++d) {
            // up_d: bit x = E[y-1][x+d] for this word and the previous one (a contact run may continue from the previous word)
            const uint32_t up = d == 0 ? u : d < 0 ? ((u << 1) | (ul >> 31)) : ((u >> 1) | (ur << 31));
            const uint32_t upp = d == 0 ? ul : d < 0 ? ((ul << 1) | (ull >> 31)) : ((ul >> 1) | (u << 31));
            uint32_t cs = lvb::run_starts(e & up, ep & upp);
            while(cs) {
                const int b = __ffs(cs) - 1; cs &= cs - 1;
                int wb = wi, bb = b + d;                       // the touched pixel of row y-1
                if(bb < 0) { bb = 31; --wb; } else if(bb > 31) { bb = 0; ++wb; }
                lvb::uf_union(A.parent, euf_run_id(A, y - 1, wb, bb), euf_run_id(A, y, wi, b));
            }
        }
    }
}
/// per row (one warp): the runs whose root is node 0 become edges (only their "maybe" pixels change in the byte mask)
__global__ void __launch_bounds__(256) edge_uf_apply(const EdgeUF A) {
    const int y = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if(y >= A.H) return;
    const int nchunks = (A.WW + 31) >> 5;
    uint32_t carry = 0;
    for(int k = 0; k < nchunks; ++k) {
        const int wi = k * 32 + lane;
        const uint32_t m = euf_word(A.E, y, wi, A.WW);
        uint32_t st = lvb::run_starts(m, euf_word(A.E, y, wi - 1, A.WW)), rs = 0;
        if(wi < A.WW) {
            const uint32_t base = 1u + (uint32_t)y * A.RS + A.rankbase[(size_t)y * A.WW + wi];
            uint32_t r = 0;
            while(st) { const int b = __ffs(st) - 1; st &= st - 1; if(lvb::uf_find(A.parent, base + r) == 0u) rs |= 1u << b; ++r; }
        }
        const uint32_t reach = lvb::fill_up_chunk(m, rs, carry);
        if(wi < A.WW) {
            uint32_t turn = reach & ~A.S[(size_t)y * A.WW + wi];   // "maybe" pixels of seeded components
            while(turn) { const int b = __ffs(turn) - 1; turn &= turn - 1; A.mask[(size_t)y * A.W + (size_t)wi * 32 + b] = EDGE_YES; }
        }
    }
}

/// the output of one threshold pass (:374: 255 where the mask holds 2), the 2 -> 3 relabel of the two persisted rows, and for apply()
/// (:412-433) the running confidence sum (+16 per threshold that finds an edge, saturated)
__global__ void __launch_bounds__(256) edge_output_kernel(uchar* mask, int W, int H, uchar* out, int accumulate) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if(x >= W || y >= H) return;
    const size_t i = (size_t)y * W + x;
    const uchar v = mask[i];
    const bool on = v >= EDGE_YES;
    if(y >= H - 2 && v == EDGE_YES) mask[i] = EDGE_STALE_YES;
    if(accumulate) { const unsigned a = out[i] + (on ? 16u : 0u); out[i] = (uchar)(a > 255u ? 255u : a); }
    else out[i] = on ? 255 : 0;
}

/// cv::normalize(confidence, confidence, 0, UCHAR_MAX, NORM_MINMAX) of apply() when the detector was built with bNormalizeOutput
/// (EdgeDetectorLBSP.cpp:431-432): min / max of the map (mm[0] starts at 255, mm[1] at 0) ...
__global__ void __launch_bounds__(256) edge_minmax_kernel(const uchar* out, size_t n, unsigned* mm) {
    unsigned lo = 255u, hi = 0u;
    for(size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) { const unsigned v = out[i]; lo = min(lo, v); hi = max(hi, v); }
#pragma unroll
    for(int o = 16; o > 0; o >>= 1) { lo = min(lo, __shfl_xor_sync(0xFFFFFFFFu, lo, o)); hi = max(hi, __shfl_xor_sync(0xFFFFFFFFu, hi, o)); }
    if((threadIdx.x & 31) == 0) { atomicMin(&mm[0], lo); atomicMax(&mm[1], hi); }
}
/// ... then OpenCV's arithmetic: scale and shift in double, the conversion in float (multiply, add, round half to even, saturate);
/// the same restatement as the oracle's, which the CPU tests pin bit-exactly against cv2 4.13
__global__ void __launch_bounds__(256) edge_normalize_kernel(uchar* out, size_t n, const unsigned* mm) {
    const double smin = (double)mm[0], smax = (double)mm[1];
    const double scale = 255.0 * (smax - smin > 2.220446049250313e-16 ? 1. / (smax - smin) : 0.);
    const float a = (float)scale, b = (float)(0.0 - smin * scale);
    for(size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int r = __float2int_rn(__fadd_rn(__fmul_rn((float)out[i], a), b));
        out[i] = (uchar)(r < 0 ? 0 : r > 255 ? 255 : r);
    }
}

} // namespace lvb_edge
