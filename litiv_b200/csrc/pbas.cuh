// litiv_b200 — PBAS per-frame kernels (replaces BackgroundSubtractorPBAS_1ch / _3ch ::initialize / apply / getBackgroundImage,
// reference video/src/BackgroundSubtractorPBAS.cpp:37-54, 60-271, 284-496; shipped switches: SELF_DIFFUSION 1, R2_ACCELERATION 0,
// ADVANCED_MORPH_OPS 0).
//
// Model: [N][H][Wp] records of (colour, gradient magnitude): 3 channels uint2 = (B|G<<8|R<<16, gB|gG<<8|gR<<16), 1 channel uchar2.
// Feedback maps R(x), T(x), mean-min-distance: three f32 planes [H][Wp]. A frame is
//   phase A  (32x8 tile, 1 thread / pixel): the input tile + 2-px halo goes to shared memory once; GaussianBlur 3x3 -> Scharr x/y ->
//            |.| -> average are evaluated there (the reference's five full-frame OpenCV passes, PBAS.cpp:125-134); scan; R/T/mean
//            feedback; own-slot write; neighbour intent; raw mask bits by ballot; frame sums of gradient distance / bad samples
//            (CTA-reduced, one 64-bit atomic each; the last CTA folds them into next frame's m_fFormerMeanGradDist)
//            The queued "self-diffusion" writes of the PREVIOUS frame are applied first (a pixel whose 3x3 neighbourhood aimed at it
//            stores ITS OWN colour / gradient of that frame, :190-191), as in vibe_phaseA.
//   phase B  the same gather, standalone, for whoever needs the model between frames (state export, getBackgroundImage)
//   pp_median (postproc.cuh) 9x9 on the bit-packed raw mask -> output mask bytes (:269 / :494)
#pragma once
#include "vibe.cuh"

namespace lvb {

struct PbasCtl { unsigned long long grad_sum; unsigned long long bad; float former; uint32_t ticket; };

struct PbasArgs {
    int W, H, Wp, WW;
    int N, REQ;
    float thr0;                // (float)nInitColorDistThreshold
    const uchar* img; size_t ipitch; int in_ch;
    void* bg; size_t plane;
    float *R, *T, *meanmin;    // [H][Wp]
    void* grad;                // [H][Wp] packed gradient of the current frame (u32 / u8), written for every pixel
    void* col;                 // [H][Wp] packed colour of the current frame (what a self-diffusion write of this frame stores)
    ushort* intents;
    const ushort* prev_intents; const void* prev_col; const void* prev_grad; // previous frame's planes, still to be applied (null: none)
    uint32_t* raw_bits;        // [H][WW]
    PbasCtl* ctl;
    uint32_t frame; uint64_t seed; uint32_t lr_override; // 0: use ceil(T(x))
    uint32_t n_magic;
    const uint32_t* magic;     // [257] floor(2^32 / n)
    unsigned long long* stats; // null or [scanned, writes, fg]
};

__device__ __forceinline__ int reflect101(int p, int len) { // cv::borderInterpolate(BORDER_REFLECT_101)
    if(len == 1) return 0;
    while(p < 0 || p >= len) { if(p < 0) p = -p; if(p >= len) p = 2 * len - 2 - p; }
    return p;
}

constexpr int PB_IN_W = 36, PB_IN_H = 12, PB_BL_W = 34, PB_BL_H = 10;

/// gradient magnitude image of the reference (PBAS.cpp:125-134) for the CTA's 32x8 tile, evaluated in shared memory.
/// Returns the packed gradient of the calling thread's pixel (0 outside the image).
template<int CH>
__device__ __forceinline__ uint32_t pbas_tile_gradient(const uchar* img, size_t ipitch, int in_ch, int W, int H, int x0, int y0,
                                                       uint32_t (*s_in)[PB_IN_W], uint32_t (*s_bl)[PB_BL_W + 2], uint32_t& cur_out) {
    const int tid = threadIdx.y * 32 + threadIdx.x;
    // 1. input tile + 2-px halo, border pixels mirrored (reflect 101): every entry is a real pixel
    for(int i = tid; i < 64 * PB_IN_H; i += 256) { // 64 slots per row, the last 28 idle: no division
        const int r = i >> 6, c = i & 63;
        if(c >= PB_IN_W) continue;
        const int gx = reflect101(min(x0 - 2 + c, W + 1), W), gy = reflect101(min(y0 - 2 + r, H + 1), H);
        s_in[r][c] = (uint32_t)vibe_load_pixel<CH>(img, ipitch, in_ch, gx, gy);
    }
    __syncthreads();
    cur_out = s_in[threadIdx.y + 2][threadIdx.x + 2];
    // 2. blurred tile + 1-px halo: [1 2 1]x[1 2 1] / 16, round half up (OpenCV's fixed-point GaussianBlur for 8-bit images).
    //    Positions outside the image take the blurred value of the mirrored position (the Scharr pass mirrors the BLURRED image).
    for(int i = tid; i < 64 * PB_BL_H; i += 256) {
        const int r = i >> 6, c = i & 63;
        const int gx = x0 - 1 + c, gy = y0 - 1 + r;
        if(c >= PB_BL_W || gx < 0 || gx >= W || gy < 0 || gy >= H) continue;
        // channels 0 / 2 and channel 1 in the 16-bit lanes of two words: a weighted sum is at most 16 * 255
        uint32_t acc02 = 0x00080008u, acc1 = 0x00000008u;
#pragma unroll
        for(int dy = 0; dy < 3; ++dy)
#pragma unroll
            for(int dx = 0; dx < 3; ++dx) {
                const uint32_t v = s_in[r + dy][c + dx];
                const uint32_t w = (dy == 1 ? 2u : 1u) * (dx == 1 ? 2u : 1u);
                acc02 += w * (v & 0x00FF00FFu);
                if(CH == 3) acc1 += w * ((v >> 8) & 0xFFu);
            }
        const uint32_t o = ((acc02 >> 4) & 0x00FF00FFu) | (((acc1 >> 4) & 0xFFu) << 8);
        s_bl[r][c] = o;
    }
    __syncthreads();
    for(int i = tid; i < 64 * PB_BL_H; i += 256) {
        const int r = i >> 6, c = i & 63;
        const int gx = x0 - 1 + c, gy = y0 - 1 + r;
        if(c >= PB_BL_W || (gx >= 0 && gx < W && gy >= 0 && gy < H)) continue;
        if(gx < -1 || gx > W || gy < -1 || gy > H) continue;          // never read by a pixel of the image
        const int mx = reflect101(gx, W) - (x0 - 1), my = reflect101(gy, H) - (y0 - 1);
        if(mx >= 0 && mx < PB_BL_W && my >= 0 && my < PB_BL_H) s_bl[r][c] = s_bl[my][mx];
    }
    __syncthreads();
    // 3. Scharr x / y (16S), convertScaleAbs (saturate |v|), addWeighted(0.5, 0.5) with round-half-even
    const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
    if(x >= W || y >= H) return 0u;
    const int r = threadIdx.y + 1, c = threadIdx.x + 1;
    const uint32_t a00 = s_bl[r - 1][c - 1], a01 = s_bl[r - 1][c], a02 = s_bl[r - 1][c + 1];
    const uint32_t a10 = s_bl[r][c - 1], a12 = s_bl[r][c + 1];
    const uint32_t a20 = s_bl[r + 1][c - 1], a21 = s_bl[r + 1][c], a22 = s_bl[r + 1][c + 1];
    // positive and negative halves of each Scharr sum in 16-bit lanes (each at most 16 * 255), then per-lane |P - M|
    auto lanes02 = [](uint32_t v) { return v & 0x00FF00FFu; };
    auto lane1 = [](uint32_t v) { return (v >> 8) & 0xFFu; };
    auto mag = [](uint32_t px, uint32_t mx, uint32_t py, uint32_t my) { // one 16-bit lane each
        const uint32_t t = min((uint32_t)abs((int)px - (int)mx), 255u) + min((uint32_t)abs((int)py - (int)my), 255u);
        return (t >> 1) + ((t & 1u) & ((t >> 1) & 1u));
    };
    const uint32_t xp = 3u * lanes02(a02) + 10u * lanes02(a12) + 3u * lanes02(a22), xm = 3u * lanes02(a00) + 10u * lanes02(a10) + 3u * lanes02(a20);
    const uint32_t yp = 3u * lanes02(a20) + 10u * lanes02(a21) + 3u * lanes02(a22), ym = 3u * lanes02(a00) + 10u * lanes02(a01) + 3u * lanes02(a02);
    uint32_t g = mag(xp & 0xFFFFu, xm & 0xFFFFu, yp & 0xFFFFu, ym & 0xFFFFu);
    if(CH == 3) {
        g |= mag(xp >> 16, xm >> 16, yp >> 16, ym >> 16) << 16;
        const uint32_t xp1 = 3u * lane1(a02) + 10u * lane1(a12) + 3u * lane1(a22), xm1 = 3u * lane1(a00) + 10u * lane1(a10) + 3u * lane1(a20);
        const uint32_t yp1 = 3u * lane1(a20) + 10u * lane1(a21) + 3u * lane1(a22), ym1 = 3u * lane1(a00) + 10u * lane1(a01) + 3u * lane1(a02);
        g |= mag(xp1, xm1, yp1, ym1) << 8;
    }
    return g;
}

template<int CH> struct PbasRec;
template<> struct PbasRec<1> { typedef uchar2 T; };
template<> struct PbasRec<3> { typedef uint2 T; };
__device__ __forceinline__ uint32_t pbas_col(const uchar2& r) { return r.x; }
__device__ __forceinline__ uint32_t pbas_grd(const uchar2& r) { return r.y; }
__device__ __forceinline__ uint32_t pbas_col(const uint2& r) { return r.x; }
__device__ __forceinline__ uint32_t pbas_grd(const uint2& r) { return r.y; }
template<int CH> __device__ __forceinline__ typename PbasRec<CH>::T pbas_rec(uint32_t c, uint32_t g) {
    if constexpr (CH == 1) return make_uchar2((uchar)c, (uchar)g); else return make_uint2(c, g);
}

/// lv::L2dist<3,uchar>: squares summed in uint16 (wraps), float sqrt (utils/math.hpp:391-397)
__device__ __forceinline__ float pbas_l2dist3(uint32_t a, uint32_t b) {
    const uint32_t ad = __vabsdiffu4(a, b);                 // the packed values keep byte 3 zero
    return __fsqrt_rn((float)(__dp4a(ad, ad, 0u) & 0xFFFFu)); // d0^2 + d1^2 + d2^2 in one IDP.4A
}

#ifndef PBAS_MIN_BLOCKS
#define PBAS_MIN_BLOCKS 6
#endif
template<int CH>
__global__ void __launch_bounds__(256, PBAS_MIN_BLOCKS) pbas_phaseA(const PbasArgs A) {
    typedef typename PbasRec<CH>::T Rec;
    __shared__ uint32_t s_in[PB_IN_H][PB_IN_W];
    __shared__ uint32_t s_bl[PB_BL_H][PB_BL_W + 2];
    __shared__ unsigned long long s_sum;
    __shared__ uint32_t s_cnt[4]; // bad | scanned | writes | fg
    __shared__ float s_former;
    const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 8;
    const int tid = threadIdx.y * 32 + threadIdx.x;
    const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
    const bool in_img = x < A.W && y < A.H;
    const size_t pix = (size_t)y * A.Wp + x;
    if(tid == 0) { s_sum = 0ull; s_former = A.ctl->former; }
    if(tid < 4) s_cnt[tid] = 0;
    // state and the first two samples are in flight while the tile is staged and the gradient evaluated
    float R = 1.f, T = 2.f, mm = 0.f;
    Rec v0 = Rec(), v1 = Rec();
    const Rec* bgr = (const Rec*)A.bg + pix;
    const uint32_t N = (uint32_t)A.N, REQ = (uint32_t)A.REQ;
    if(in_img) {
        R = A.R[pix]; T = A.T[pix]; mm = A.meanmin[pix];
        v0 = bgr[0];
        if(N > 1u) v1 = bgr[A.plane];
    }
    if(A.prev_intents != nullptr) { // neighbour writes queued by the previous frame (see vibe_phaseA)
        __shared__ ushort s_int[10][36];
        __shared__ uint32_t s_hits[8][32];
        uint32_t hits = vibe_stage_hits(A.prev_intents, A.W, A.H, A.Wp, x0, y0, s_int, s_hits);
        if(in_img && hits) { // self-diffusion: every hit stores this pixel's own colour / gradient of the previous frame
            const uint32_t c = CH == 1 ? (uint32_t)((const uchar*)A.prev_col)[pix] : ((const uint32_t*)A.prev_col)[pix];
            const uint32_t g = CH == 1 ? (uint32_t)((const uchar*)A.prev_grad)[pix] : ((const uint32_t*)A.prev_grad)[pix];
            const Rec own = pbas_rec<CH>(c, g);
            while(hits) {
                const int i = __ffs(hits) - 1, dy = i / 3 - 1, dx = i - (i / 3) * 3 - 1;
                hits &= hits - 1u;
                const uint32_t slot = s_int[threadIdx.y + 1 + dy][threadIdx.x + 1 + dx] & 0xFFu;
                ((Rec*)A.bg)[(size_t)slot * A.plane + pix] = own;
                if(slot == 0u) v0 = own;
                if(slot == 1u) v1 = own;
            }
        }
    }
    uint32_t cur;
    const uint32_t cg = pbas_tile_gradient<CH>(A.img, A.ipitch, A.in_ch, A.W, A.H, x0, y0, s_in, s_bl, cur);

    bool is_fg = false;
    uint32_t scanned = 0, writes = 0, bad = 0;
    unsigned long long gsum = 0ull;
    if(in_img) {
        const float grad_w = __fdiv_rn(10.0f, s_former);      // BGSPBAS_GRAD_WEIGHT_ALPHA / m_fFormerMeanGradDist
        const float thr = __fmul_rn(R, A.thr0);
        float min_dist = 255.0f;
        uint32_t good = 0, s = 0;
        auto test = [&](const Rec& v) {
            float sum; unsigned long long gfix;
            if constexpr (CH == 1) {
                const uint32_t cd = (uint32_t)abs((int)cur - (int)pbas_col(v)), gd = (uint32_t)abs((int)cg - (int)pbas_grd(v));
                sum = fminf(__fadd_rn(__fmul_rn(grad_w, (float)gd), (float)cd), 255.0f);
                gfix = gd;
            } else {
                const float cd = pbas_l2dist3(cur, pbas_col(v)), gd = pbas_l2dist3(cg, pbas_grd(v));
                sum = fminf(__fadd_rn(__fmul_rn(grad_w, gd), cd), 255.0f);
                gfix = __float2ull_rn(__fmul_rn(gd, 65536.0f)); // 2^-16 fixed point: the frame sum is order-independent
            }
            if(sum <= thr) { if(min_dist > sum) min_dist = sum; ++good; }
            else { gsum += gfix; ++bad; }
            ++s;
        };
        if(good < REQ && s < N) test(v0);
        if(good < REQ && s < N) test(v1);
        while(good < REQ && s < N) {
            const bool two = s + 1u < N;
            v0 = bgr[(size_t)s * A.plane];
            if(two) v1 = bgr[(size_t)(s + 1u) * A.plane];
            test(v0);
            if(good < REQ && two) test(v1);
        }
        scanned = s;
        const float fN = (float)N;
        mm = __fdiv_rn(__fadd_rn(__fmul_rn(mm, (float)(N - 1u)), __fdiv_rn(min_dist, 255.0f)), fN);
        const float tden = __fadd_rn(__fmul_rn(mm, 255.0f), 1.0f);
        uint32_t intent = VIBE_NO_INTENT;
        if(good < REQ) {
            is_fg = true;
            T = __fadd_rn(T, __fdiv_rn(1.0f, tden));
            if(T > 200.0f) T = 200.0f;
        } else {
            const uint32_t lr = A.lr_override ? A.lr_override : (uint32_t)ceilf(T);
            const uint32_t lrm = lr <= 256u ? A.magic[lr] : (lr <= 1u ? 0xFFFFFFFFu : (uint32_t)(0x100000000ull / lr));
            const uint4 rnd = philox_block(A.seed, A.frame, (uint32_t)(y * A.W + x), 0, DOM_APPLY);
            if(fast_mod(rnd.x, lr, lrm) == 0u) {
                ((Rec*)A.bg)[(size_t)fast_mod(rnd.y, N, A.n_magic) * A.plane + pix] = pbas_rec<CH>(cur, cg);
                ++writes;
            }
            if(fast_mod(rnd.z, lr, lrm) == 0u) {
                int dx, dy;
                neighbor_offset(true, rnd.w, dx, dy);
                const int nx = clampi(x + dx, 0, A.W - 1), ny = clampi(y + dy, 0, A.H - 1);
                intent = (uint32_t)(((ny - y + 1) * 3 + (nx - x + 1)) << 8) | fast_mod(fast_div(rnd.y, N, A.n_magic), N, A.n_magic);
                ++writes;
            }
            T = __fsub_rn(T, __fdiv_rn(0.05f, tden));
            if(T < 2.0f) T = 2.0f;
        }
        if(R < __fadd_rn(__fadd_rn(0.6f, __fmul_rn(mm, 5.0f)), 0.0f)) { if(R < 99.0f) R = __fmul_rn(R, 1.05f); }
        else if(R > 0.6f) R = __fmul_rn(R, 0.95f);
        A.R[pix] = R; A.T[pix] = T; A.meanmin[pix] = mm;
        A.intents[pix] = (ushort)intent;
        if constexpr (CH == 1) { ((uchar*)A.grad)[pix] = (uchar)cg; ((uchar*)A.col)[pix] = (uchar)cur; }
        else { ((uint32_t*)A.grad)[pix] = cg; ((uint32_t*)A.col)[pix] = cur; }
    }
    const uint32_t b_raw = __ballot_sync(0xFFFFFFFFu, is_fg);
    if(threadIdx.x == 0 && y < A.H && (x >> 5) < A.WW) A.raw_bits[y * A.WW + (x >> 5)] = b_raw;
    // frame sums: warp shuffle -> shared -> one global atomic per CTA and counter
#pragma unroll
    for(int o = 16; o > 0; o >>= 1) {
        gsum += __shfl_xor_sync(0xFFFFFFFFu, gsum, o); bad += __shfl_xor_sync(0xFFFFFFFFu, bad, o);
        scanned += __shfl_xor_sync(0xFFFFFFFFu, scanned, o); writes += __shfl_xor_sync(0xFFFFFFFFu, writes, o);
    }
    if(threadIdx.x == 0) {
        if(gsum) atomicAdd(&s_sum, gsum);
        if(bad) atomicAdd(&s_cnt[0], bad);
        if(A.stats) { atomicAdd(&s_cnt[1], scanned); atomicAdd(&s_cnt[2], writes); atomicAdd(&s_cnt[3], __popc(b_raw)); }
    }
    __syncthreads();
    if(tid == 0) {
        if(s_sum) atomicAdd(&A.ctl->grad_sum, s_sum);
        if(s_cnt[0]) atomicAdd(&A.ctl->bad, (unsigned long long)s_cnt[0]);
        if(A.stats) for(int i = 0; i < 3; ++i) if(s_cnt[1 + i]) atomicAdd(&A.stats[i], (unsigned long long)s_cnt[1 + i]);
        __threadfence();
        // the last CTA of the frame folds the sums into next frame's m_fFormerMeanGradDist (PBAS.cpp:224 / :456) and clears them
        if(atomicAdd(&A.ctl->ticket, 1u) == gridDim.x * gridDim.y - 1u) {
            __threadfence();
            const unsigned long long gs = atomicAdd(&A.ctl->grad_sum, 0ull), bd = atomicAdd(&A.ctl->bad, 0ull) + 1ull; // count starts at 1
            const float tot = CH == 1 ? __ull2float_rn(gs) : __double2float_rn(__dmul_rn(__ull2double_rn(gs), 1.0 / 65536.0));
            A.ctl->former = fmaxf(__fdiv_rn(tot, __ull2float_rn(bd)), 20.0f);
            A.ctl->grad_sum = 0ull; A.ctl->bad = 0ull; A.ctl->ticket = 0u;
        }
    }
}

/// queued neighbour writes with BGSPBAS_USE_SELF_DIFFUSION (PBAS.cpp:186-191 / :420-425): the TARGET stores its own colour and
/// gradient of that frame in the drawn slot, so colliding writes are idempotent and only "did a neighbour aim at me, with which
/// slot" matters. Standalone form of the gather at the top of pbas_phaseA; reads A.prev_intents / prev_col / prev_grad.
template<int CH>
__global__ void __launch_bounds__(256) pbas_phaseB(const PbasArgs A) {
    typedef typename PbasRec<CH>::T Rec;
    __shared__ ushort s_int[10][36];
    __shared__ uint32_t s_hits[8][32];
    const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 8;
    uint32_t hits = vibe_stage_hits(A.prev_intents, A.W, A.H, A.Wp, x0, y0, s_int, s_hits);
    const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
    if(x >= A.W || y >= A.H || !hits) return;
    const size_t pix = (size_t)y * A.Wp + x;
    const uint32_t c = CH == 1 ? (uint32_t)((const uchar*)A.prev_col)[pix] : ((const uint32_t*)A.prev_col)[pix];
    const uint32_t g = CH == 1 ? (uint32_t)((const uchar*)A.prev_grad)[pix] : ((const uint32_t*)A.prev_grad)[pix];
    const Rec own = pbas_rec<CH>(c, g);
    while(hits) {
        const int i = __ffs(hits) - 1, dy = i / 3 - 1, dx = i - (i / 3) * 3 - 1;
        hits &= hits - 1u;
        ((Rec*)A.bg)[(size_t)(s_int[threadIdx.y + 1 + dy][threadIdx.x + 1 + dx] & 0xFFu) * A.plane + pix] = own;
    }
}

/// gradient image alone (initialize: PBAS.cpp:80-89 / :301-310)
template<int CH>
__global__ void __launch_bounds__(256) pbas_grad_kernel(const PbasArgs A) {
    __shared__ uint32_t s_in[PB_IN_H][PB_IN_W];
    __shared__ uint32_t s_bl[PB_BL_H][PB_BL_W + 2];
    const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 8;
    uint32_t cur;
    const uint32_t g = pbas_tile_gradient<CH>(A.img, A.ipitch, A.in_ch, A.W, A.H, x0, y0, s_in, s_bl, cur);
    const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
    if(x >= A.W || y >= A.H) return;
    const size_t pix = (size_t)y * A.Wp + x;
    if constexpr (CH == 1) ((uchar*)A.grad)[pix] = (uchar)g; else ((uint32_t*)A.grad)[pix] = g;
}

/// initialize (PBAS.cpp:91-108 / :312-325): colour and gradient of sample s from the same 7x7-distributed neighbour (border 0)
template<int CH>
__global__ void __launch_bounds__(256) pbas_init_kernel(const PbasArgs A, float t0) {
    typedef typename PbasRec<CH>::T Rec;
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if(x >= A.W || y >= A.H) return;
    const size_t pix = (size_t)y * A.Wp + x;
    const uint32_t p = (uint32_t)(y * A.W + x);
    for(int s = 0; s < A.N; ++s) {
        const uint32_t rnd = philox_draw(A.seed, 0u, p, (uint32_t)s, DOM_REFRESH);
        int r = 1 + (int)(rnd % 512u), i = 0;
        for(; i < 49; ++i) { r -= c_pat7[i]; if(r <= 0) break; }
        if(i > 48) i = 48;
        const int sx = clampi(x + (i % 7) - 3, 0, A.W - 1), sy = clampi(y + (i / 7) - 3, 0, A.H - 1);
        const size_t q = (size_t)sy * A.Wp + sx;
        const uint32_t c = (uint32_t)vibe_load_pixel<CH>(A.img, A.ipitch, A.in_ch, sx, sy);
        const uint32_t g = CH == 1 ? (uint32_t)((const uchar*)A.grad)[q] : ((const uint32_t*)A.grad)[q];
        ((Rec*)A.bg)[(size_t)s * A.plane + pix] = pbas_rec<CH>(c, g);
    }
    A.R[pix] = 1.0f; A.T[pix] = t0; A.meanmin[pix] = 0.0f;
}

/// getBackgroundImage (PBAS.cpp:37-54)
template<int CH>
__global__ void __launch_bounds__(256) pbas_background_kernel(const PbasArgs A, uchar* out) {
    typedef typename PbasRec<CH>::T Rec;
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if(x >= A.W || y >= A.H) return;
    const size_t pix = (size_t)y * A.Wp + x;
    float acc[CH];
#pragma unroll
    for(int c = 0; c < CH; ++c) acc[c] = 0.f;
    for(int s = 0; s < A.N; ++s) {
        const uint32_t v = pbas_col(((const Rec*)A.bg)[(size_t)s * A.plane + pix]);
#pragma unroll
        for(int c = 0; c < CH; ++c) acc[c] = __fadd_rn(acc[c], __fdiv_rn((float)((v >> (8 * c)) & 0xFFu), (float)A.N));
    }
#pragma unroll
    for(int c = 0; c < CH; ++c) out[((size_t)y * A.W + x) * CH + c] = (uchar)fminf(fmaxf(rintf(acc[c]), 0.f), 255.f);
}

/// state export / import of the model: which = 0 colour, 1 gradient; reference layout [N][H][W][C]
template<int CH>
__global__ void __launch_bounds__(256) pbas_model_copy_kernel(const PbasArgs A, uchar* ref_layout, int which, int to_device) {
    typedef typename PbasRec<CH>::T Rec;
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y, s = blockIdx.z;
    if(x >= A.W || y >= A.H) return;
    Rec* d = (Rec*)A.bg + (size_t)s * A.plane + (size_t)y * A.Wp + x;
    uchar* r = ref_layout + (((size_t)s * A.H + y) * A.W + x) * CH;
    Rec rec = *d;
    if(to_device) {
        uint32_t v = 0;
#pragma unroll
        for(int c = 0; c < CH; ++c) v |= (uint32_t)r[c] << (8 * c);
        *d = which == 0 ? pbas_rec<CH>(v, pbas_grd(rec)) : pbas_rec<CH>(pbas_col(rec), v);
    } else {
        const uint32_t v = which == 0 ? pbas_col(rec) : pbas_grd(rec);
#pragma unroll
        for(int c = 0; c < CH; ++c) r[c] = (uchar)(v >> (8 * c));
    }
}

} // namespace lvb
