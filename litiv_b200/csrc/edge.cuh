// litiv_b200 — kernels of the LBSP edge detector (reference imgproc/src/EdgeDetectorLBSP.cpp:47-433). The per-pixel
// bodies live in edge_px.cuh (shared with the CPU emulation of the tests); the per-level LBSP gradient is lbsp_gradient_kernel of
// lobster.cuh (TMA-staged tile + 5x5 halo). All maps are HBM-resident per detector object. Algorithmic bytes of one apply_threshold
// call per full-resolution pixel, three levels (sizes 1 + 1/4 + 1/16 = 1.31): image read C and 4-byte map written, re-read and rewritten by
// the combination (C + 13 per level pixel, + C for writing the two coarser images), the suppression's map read 4 and mask write 1 (its 5x5
// window is served by L1 / L2), one mask read per flood sweep, mask read + result write 2: about 33 B/px for RGB with four sweeps
// (tools/bench_edge.py). The detector is launch bound at CDnet sizes (16 launches and one flag read-back per call).
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

/// V_l = combine(own gradient, coarser V or the initial value); `own` and `out` may alias (one thread per pixel, read before write)
__global__ void __launch_bounds__(256) edge_combine_kernel(const uchar4* own, int Wl, int Hl, const uchar4* coarse, int Wc, uchar4* out) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if(x >= Wl || y >= Hl) return;
    const uchar4 c = coarse ? coarse[(size_t)(y >> 1) * Wc + (x >> 1)] : edge_init_value();
    out[(size_t)y * Wl + x] = edge_combine(own[(size_t)y * Wl + x], c);
}

/// mask rows 0 .. H-3 <- suppression class of gradient rows 2 .. H-1; rows H-2, H-1 are left as the previous call left them
__global__ void __launch_bounds__(256) edge_nms_kernel(const EdgeMaps m, unsigned lo, unsigned hi, uchar* mask) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if(x >= m.W || y >= m.H - 2) return;
    mask[(size_t)y * m.W + x] = edge_mask_value(m, y, x, lo, hi, mask);
}

/// hysteresis (:353-372): one relaxation sweep. Every CTA iterates its 32x16 tile (+1 halo) in shared memory until the tile is stable
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
