// litiv_b200 — ViBe per-frame kernels (replaces BackgroundSubtractorViBe_1ch / _3ch ::initialize / apply / getBackgroundImage,
// reference video/src/BackgroundSubtractorViBe.cpp:32-49, 58-110, 115-191): the colour-only ancestor of the LOBSTER scan.
//
// Model: [N][H][Wp] of one packed colour per sample (3 channels: B | G<<8 | R<<16 in a u32; 1 channel: one byte), sample-major so
// that a warp reads 128 (32) contiguous bytes per sample row. No ROI, no border, no descriptors, no post-processing: a frame is
//   ONE kernel (1 thread / pixel): the neighbour writes queued by the PREVIOUS frame are applied first (every pixel gathers the
//   intents of its 3x3 neighbourhood aimed at it from a shared-memory tile, in raster order of the source = last writer wins, and
//   copies the source's colour of that frame into the drawn slot); then scan -> mask byte; own-slot write at once; this frame's
//   neighbour write queued as an intent word + the pixel's colour. Anything that needs the model between frames (export,
//   getBackgroundImage) applies the pending writes with the standalone vibe_phaseB.
// HBM bound by construction: ~(C + 1 + 2 + 2) + s*Cpacked bytes per pixel for s scanned samples and a handful of instructions each.
#pragma once
#include "common.cuh"

namespace lvb {

constexpr uint32_t VIBE_NO_INTENT = 0xFFFFu; // valid intents are (code 0..8) << 8 | slot

struct VibeArgs {
    int W, H, Wp;
    int N, REQ;
    uint32_t thr;              // 1 channel: nColorDistThreshold (L1 < thr) ; 3 channels: (nColorDistThreshold*3)^2 (squared L2 < thr)
    const uchar* img; size_t ipitch; int in_ch; // in_ch == 1 with a 3-channel model: cvtColor(GRAY2BGR) on the fly (ViBe.cpp:121-124)
    void* bg; size_t plane;    // plane = H*Wp samples
    ushort* intents;           // [H][Wp] intent words this frame writes
    void* nbcol;               // [H][Wp] colour of the pixels that queued a neighbour write this frame (written only there)
    const ushort* prev_intents; const void* prev_nbcol; // previous frame's planes, still to be applied (null: nothing pending)
    uchar* mask; size_t mpitch;
    uint32_t frame; uint64_t seed; uint32_t lr; // lr = ceil(learningRate), 0xFFFFFFFF for +inf
    uint32_t lr_magic, n_magic;  // floor(2^32 / lr), floor(2^32 / N) for fast_mod / fast_div
    unsigned long long* stats; // null or [scanned, writes, fg]
};

template<int CH> struct VibeCol;
template<> struct VibeCol<1> { typedef uchar T; };
template<> struct VibeCol<3> { typedef uint32_t T; };

template<int CH>
__device__ __forceinline__ typename VibeCol<CH>::T vibe_load_pixel(const uchar* img, size_t ipitch, int in_ch, int x, int y) {
    const uchar* p = img + (size_t)y * ipitch;
    if constexpr (CH == 1) return p[x];
    else {
        if(in_ch == 1) { const uint32_t g = p[x]; return g | (g << 8) | (g << 16); }
        p += (size_t)x * 3;
        return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16);
    }
}

/// ViBe.cpp:93 (L1dist < thr) / :167-171 (L2dist < thr*3). lv::L2dist<3,uchar> accumulates the squares in uint16 (utils/math.hpp:
/// 391-397, L2sqrdist :301-306): the sum wraps mod 65536 before the float square root. For an integer n < 2^16 and an integer
/// T <= 765, (float)sqrt(n) < T  <=>  n < T^2 (sqrt is correctly rounded and T - sqrt(T^2 - 1) > 6e-4 >> ulp(T)/2), so the kernel
/// compares integers; the oracle keeps the float form.
template<int CH>
__device__ __forceinline__ bool vibe_match(typename VibeCol<CH>::T cur, typename VibeCol<CH>::T b, uint32_t thr) {
    if constexpr (CH == 1) return (uint32_t)abs((int)cur - (int)b) < thr;
    else {
        const uint32_t ad = __vabsdiffu4(cur, b);           // the packed values keep byte 3 zero
        const uint32_t acc = __dp4a(ad, ad, 0u) & 0xFFFFu;  // d0^2 + d1^2 + d2^2 in one IDP.4A
        return acc < thr;
    }
}

/// Stages the intent words of a 32x8 tile + 1-px halo in shared memory ("none" outside the image) and returns, for the calling
/// thread's pixel, the mask of 3x3 neighbourhood positions whose pixel aims at it: bit (dy+1)*3 + (dx+1) for the source at
/// (x+dx, y+dy), so ascending bits = raster order of the sources. Scatter inside the CTA: the ~6 % of entries that carry an intent
/// mark their target with a shared-memory atomicOr, instead of every pixel testing its nine neighbours. Two barriers.
__device__ __forceinline__ uint32_t vibe_stage_hits(const ushort* intents, int W, int H, int Wp, int x0, int y0, ushort (*s_int)[36],
                                                    uint32_t (*s_hits)[32]) {
    const int tid = threadIdx.y * 32 + threadIdx.x;
    s_hits[threadIdx.y][threadIdx.x] = 0u;
    __syncthreads();
    auto fetch = [&](int r, int c) { // tile coordinates incl. the halo
        const int gx = x0 - 1 + c, gy = y0 - 1 + r;
        const uint32_t it = (gx >= 0 && gx < W && gy >= 0 && gy < H) ? (uint32_t)intents[(size_t)gy * Wp + gx] : VIBE_NO_INTENT;
        s_int[r][c] = (ushort)it;
        if(it != VIBE_NO_INTENT) {
            const int code = (int)(it >> 8), oy = code / 3 - 1, ox = code - (code / 3) * 3 - 1; // source -> target offset (clamped)
            const int tr = r - 1 + oy, tc = c - 1 + ox;                                         // target inside the 32x8 core?
            if(tr >= 0 && tr < 8 && tc >= 0 && tc < 32) atomicOr(&s_hits[tr][tc], 1u << ((1 - oy) * 3 + (1 - ox)));
        }
    };
    fetch(threadIdx.y + 1, threadIdx.x + 1);                          // core: one coalesced row per warp
    if(tid < 68) fetch(tid < 34 ? 0 : 9, tid < 34 ? tid : tid - 34);  // top / bottom halo rows
    else if(tid < 84) fetch(1 + ((tid - 68) & 7), tid < 76 ? 0 : 33); // left / right halo columns
    __syncthreads();
    return s_hits[threadIdx.y][threadIdx.x];
}

#ifndef VIBE_MIN_BLOCKS
#define VIBE_MIN_BLOCKS 8
#endif
template<int CH>
__global__ void __launch_bounds__(256, VIBE_MIN_BLOCKS) vibe_phaseA(const VibeArgs A) {
    typedef typename VibeCol<CH>::T Col;
    __shared__ ushort s_int[10][36];
    const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 8;
    const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
    const bool in_img = x < A.W && y < A.H;
    const bool pending = A.prev_intents != nullptr;
    uint32_t scanned = 0, writes = 0;
    bool is_fg = false;
    const size_t pix = (size_t)y * A.Wp + x;
    const Col* bgr = (const Col*)A.bg + pix;
    const uint32_t N = (uint32_t)A.N, REQ = (uint32_t)A.REQ;
    // the first two samples are (almost) always scanned: both are in flight before anything else happens
    Col v0 = Col(), v1 = Col(), cur = Col();
    if(in_img) {
        v0 = bgr[0];
        if(N > 1u) v1 = bgr[A.plane];
        cur = vibe_load_pixel<CH>(A.img, A.ipitch, A.in_ch, x, y);
    }
    if(pending) { // neighbour writes queued by the previous frame, in raster order of their source (last writer wins)
        __shared__ uint32_t s_hits[8][32];
        uint32_t hits = vibe_stage_hits(A.prev_intents, A.W, A.H, A.Wp, x0, y0, s_int, s_hits);
        if(!in_img) hits = 0u;
        while(hits) {
            const int i = __ffs(hits) - 1, dy = i / 3 - 1, dx = i - (i / 3) * 3 - 1;
            hits &= hits - 1u;
            const uint32_t slot = s_int[threadIdx.y + 1 + dy][threadIdx.x + 1 + dx] & 0xFFu;
            const Col c = ((const Col*)A.prev_nbcol)[(size_t)(y + dy) * A.Wp + (x + dx)];
            ((Col*)A.bg)[(size_t)slot * A.plane + pix] = c; // later loads of this thread see it (same thread, same address)
            if(slot == 0u) v0 = c;
            if(slot == 1u) v1 = c;
        }
    }
    if(in_img) {
        uint32_t good = 0, s = 0;
        if(good < REQ && s < N) { good += vibe_match<CH>(cur, v0, A.thr) ? 1u : 0u; ++s; }
        if(good < REQ && s < N) { good += vibe_match<CH>(cur, v1, A.thr) ? 1u : 0u; ++s; }
        // undecided after two samples (~10 % of the pixels): one more pair, then eight samples in flight per DRAM round trip. A
        // foreground pixel scans all N, and its warp (and CTA slot) waits for it: the deeper batches cut that tail's latency 4x.
        if(good < REQ && s < N) {
            const bool two = s + 1u < N;
            v0 = bgr[(size_t)s * A.plane];
            if(two) v1 = bgr[(size_t)(s + 1u) * A.plane];
            good += vibe_match<CH>(cur, v0, A.thr) ? 1u : 0u; ++s;
            if(good < REQ && two) { good += vibe_match<CH>(cur, v1, A.thr) ? 1u : 0u; ++s; }
        }
        while(good < REQ && s < N) {
            Col v[8];
            const uint32_t s0 = s;
#pragma unroll
            for(uint32_t j = 0; j < 8u; ++j) if(s0 + j < N) v[j] = bgr[(size_t)(s0 + j) * A.plane];
#pragma unroll
            for(uint32_t j = 0; j < 8u; ++j) if(good < REQ && s0 + j < N) { good += vibe_match<CH>(cur, v[j], A.thr) ? 1u : 0u; ++s; }
        }
        scanned = s;
        uint32_t intent = VIBE_NO_INTENT;
        if(good < REQ) is_fg = true;
        else { // ViBe.cpp:101-108 / :181-188
            const uint4 rnd = philox_block(A.seed, A.frame, (uint32_t)(y * A.W + x), 0, DOM_APPLY);
            if(fast_mod(rnd.x, A.lr, A.lr_magic) == 0u) {
                ((Col*)A.bg)[(size_t)fast_mod(rnd.y, N, A.n_magic) * A.plane + pix] = cur;
                ++writes;
            }
            if(fast_mod(rnd.z, A.lr, A.lr_magic) == 0u) {
                int dx, dy;
                neighbor_offset(true, rnd.w, dx, dy);
                const int nx = clampi(x + dx, 0, A.W - 1), ny = clampi(y + dy, 0, A.H - 1); // border 0 (getNeighborPosition_3x3(...,0,size))
                intent = (uint32_t)(((ny - y + 1) * 3 + (nx - x + 1)) << 8) | fast_mod(fast_div(rnd.y, N, A.n_magic), N, A.n_magic); // one Philox block per pixel
                ((Col*)A.nbcol)[pix] = cur;
                ++writes;
            }
        }
        A.intents[pix] = (ushort)intent;
        A.mask[(size_t)y * A.mpitch + x] = is_fg ? 255 : 0;
    }
    if(A.stats) { // instrumentation only: per-CTA sums in shared memory, one global atomic per counter and CTA
        __shared__ uint32_t s_cnt[3];
        const int tid = threadIdx.y * 32 + threadIdx.x;
        if(tid < 3) s_cnt[tid] = 0;
        __syncthreads();
        uint32_t sc = scanned, wr = writes, fg = is_fg ? 1u : 0u;
#pragma unroll
        for(int o = 16; o > 0; o >>= 1) { sc += __shfl_xor_sync(0xFFFFFFFFu, sc, o); wr += __shfl_xor_sync(0xFFFFFFFFu, wr, o); fg += __shfl_xor_sync(0xFFFFFFFFu, fg, o); }
        if(threadIdx.x == 0) { atomicAdd(&s_cnt[0], sc); atomicAdd(&s_cnt[1], wr); atomicAdd(&s_cnt[2], fg); }
        __syncthreads();
        if(tid < 3 && s_cnt[tid]) atomicAdd(&A.stats[tid], (unsigned long long)s_cnt[tid]);
    }
}

/// queued neighbour writes (ViBe.cpp:104-107 / :184-187), gathered per TARGET pixel so that colliding writes resolve in raster
/// order of their source without atomics. Standalone form of the gather at the top of vibe_phaseA, for whoever needs the model
/// before the next frame arrives. Reads A.prev_intents / A.prev_nbcol.
template<int CH>
__global__ void __launch_bounds__(256) vibe_phaseB(const VibeArgs A) {
    typedef typename VibeCol<CH>::T Col;
    __shared__ ushort s_int[10][36];
    __shared__ uint32_t s_hits[8][32];
    const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 8;
    uint32_t hits = vibe_stage_hits(A.prev_intents, A.W, A.H, A.Wp, x0, y0, s_int, s_hits);
    const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
    if(x >= A.W || y >= A.H) return;
    const size_t pix = (size_t)y * A.Wp + x;
    while(hits) { // ascending bits = raster order of the sources
        const int i = __ffs(hits) - 1, dy = i / 3 - 1, dx = i - (i / 3) * 3 - 1;
        hits &= hits - 1u;
        const uint32_t slot = s_int[threadIdx.y + 1 + dy][threadIdx.x + 1 + dx] & 0xFFu;
        ((Col*)A.bg)[(size_t)slot * A.plane + pix] = ((const Col*)A.prev_nbcol)[(size_t)(y + dy) * A.Wp + (x + dx)];
    }
}

/// initialize (ViBe.cpp:58-76 / :115-138): sample s of pixel p = the init image at a 7x7-Gaussian-distributed neighbour (border 0)
template<int CH>
__global__ void __launch_bounds__(256) vibe_init_kernel(const VibeArgs A) {
    typedef typename VibeCol<CH>::T Col;
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
        ((Col*)A.bg)[(size_t)s * A.plane + pix] = vibe_load_pixel<CH>(A.img, A.ipitch, A.in_ch, sx, sy);
    }
}

/// getBackgroundImage (ViBe.cpp:32-49): float mean accumulated sample by sample, round-half-even + saturate
template<int CH>
__global__ void __launch_bounds__(256) vibe_background_kernel(const VibeArgs A, uchar* out) {
    typedef typename VibeCol<CH>::T Col;
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if(x >= A.W || y >= A.H) return;
    const size_t pix = (size_t)y * A.Wp + x;
    float acc[CH];
#pragma unroll
    for(int c = 0; c < CH; ++c) acc[c] = 0.f;
    for(int s = 0; s < A.N; ++s) {
        const uint32_t v = ((const Col*)A.bg)[(size_t)s * A.plane + pix];
#pragma unroll
        for(int c = 0; c < CH; ++c) acc[c] = __fadd_rn(acc[c], __fdiv_rn((float)((v >> (8 * c)) & 0xFFu), (float)A.N));
    }
#pragma unroll
    for(int c = 0; c < CH; ++c) out[((size_t)y * A.W + x) * CH + c] = (uchar)fminf(fmaxf(rintf(acc[c]), 0.f), 255.f);
}

/// state export / import: model between the device layout and the reference's [N][H][W][C]
template<int CH>
__global__ void __launch_bounds__(256) vibe_model_copy_kernel(const VibeArgs A, uchar* ref_layout, int to_device) {
    typedef typename VibeCol<CH>::T Col;
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y, s = blockIdx.z;
    if(x >= A.W || y >= A.H) return;
    Col* d = (Col*)A.bg + (size_t)s * A.plane + (size_t)y * A.Wp + x;
    uchar* r = ref_layout + (((size_t)s * A.H + y) * A.W + x) * CH;
    if(to_device) {
        if constexpr (CH == 1) *d = r[0]; else *d = (uint32_t)r[0] | ((uint32_t)r[1] << 8) | ((uint32_t)r[2] << 16);
    } else {
        const uint32_t v = *d;
#pragma unroll
        for(int c = 0; c < CH; ++c) r[c] = (uchar)(v >> (8 * c));
    }
}

} // namespace lvb
