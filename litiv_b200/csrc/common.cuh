// litiv_b200 — device-side building blocks shared by every kernel (sm_100a only).
//   * Philox4x32-10 counter RNG (replaces libc rand())
//   * LBSP 16-bit double-cross lookup/threshold on packed bytes (features2d LBSP.hpp:193-224, 275-319)
//   * sampling patterns (utils opencv.hpp:859-966)
//   * bit-packed row helpers for the mask post-processing
//   * TMA (cp.async.bulk.tensor) + mbarrier PTX wrappers
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include <stdint.h>

namespace lvb {

typedef unsigned char uchar;
typedef unsigned short ushort;

// ---------------------------------------------------------------------------------------------
// Philox4x32-10
// ---------------------------------------------------------------------------------------------
enum { DOM_APPLY = 0, DOM_REFRESH = 1, DOM_REFRESH_START = 2, DOM_PAWCS_A = 3, DOM_PAWCS_B = 4 };

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for(int i = 0; i < 10; ++i) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u; k.y += 0xBB67AE85u;
    }
    return c;
}
/// block `blk` of 4 draws for (frame,pixel); each draw is 31 bits like libc rand()
__device__ __forceinline__ uint4 philox_block(uint64_t seed, uint32_t frame, uint32_t pixel, uint32_t blk, uint32_t domain) {
    uint4 r = philox4x32_10(make_uint4(frame, pixel, blk, domain), make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    r.x >>= 1; r.y >>= 1; r.z >>= 1; r.w >>= 1;
    return r;
}
__device__ __forceinline__ uint32_t philox_draw(uint64_t seed, uint32_t frame, uint32_t pixel, uint32_t site, uint32_t domain) {
    const uint4 r = philox_block(seed, frame, pixel, site >> 2, domain);
    const uint32_t k = site & 3;
    return k == 0 ? r.x : k == 1 ? r.y : k == 2 ? r.z : r.w;
}

// ---------------------------------------------------------------------------------------------
// LBSP on packed bytes
// ---------------------------------------------------------------------------------------------
// bit n <-> (dx,dy): LBSP.hpp:292-294
__device__ __constant__ const signed char c_lbsp_dx[16] = {-2, 2, 0, 0, -2, 2, 2, -2, 0, -1, 0, 1, -1, 1, 1, -1};
__device__ __constant__ const signed char c_lbsp_dy[16] = { 0, 0,-2, 2,  2,-2, 2, -2, 1,  0,-1, 0, -1, 1,-1,  1};

struct Lookup16 { uint32_t w[4]; }; // 16 neighbour bytes, byte n of the array = neighbour n

/// gather the 16 neighbours of channel c around (sx,sy) from an interleaved byte tile in shared memory
template<int CH>
__device__ __forceinline__ Lookup16 lbsp_lookup_smem(const uchar* tile, int pitch, int sx, int sy, int c) {
    const uchar* p = tile + sy * pitch + sx * CH + c;
    Lookup16 L;
#define LVB_NB(dx, dy) ((uint32_t)p[(dy) * pitch + (dx) * CH])
    L.w[0] = LVB_NB(-2, 0) | (LVB_NB(2, 0) << 8) | (LVB_NB(0, -2) << 16) | (LVB_NB(0, 2) << 24);
    L.w[1] = LVB_NB(-2, 2) | (LVB_NB(2, -2) << 8) | (LVB_NB(2, 2) << 16) | (LVB_NB(-2, -2) << 24);
    L.w[2] = LVB_NB(0, 1) | (LVB_NB(-1, 0) << 8) | (LVB_NB(0, -1) << 16) | (LVB_NB(1, 0) << 24);
    L.w[3] = LVB_NB(-1, -1) | (LVB_NB(1, 1) << 8) | (LVB_NB(1, -1) << 16) | (LVB_NB(-1, 1) << 24);
#undef LVB_NB
    return L;
}

/// 5x5 neighbourhood of one pixel as five realigned row windows (bytes dxi*CH + c, dxi = dx+2), fetched with aligned
/// 32-bit shared loads + funnel shifts instead of 16*CH byte loads; `o` = byte offset of pixel (sx-2) inside a tile row
template<int CH> struct Window5 { uint32_t a[5][(5 * CH + 3) / 4]; };
template<int CH>
__device__ __forceinline__ Window5<CH> lbsp_window_smem(const uchar* tile, int pitch, int sy, int o) {
    constexpr int NW = (5 * CH + 3) / 4;
    Window5<CH> Wn;
    const int sh = (o & 3) * 8;
#pragma unroll
    for(int r = 0; r < 5; ++r) {
        const uint32_t* rp = (const uint32_t*)(tile + (sy - 2 + r) * pitch + (o & ~3));
        uint32_t w[NW + 1];
#pragma unroll
        for(int i = 0; i <= NW; ++i) w[i] = rp[i];
#pragma unroll
        for(int i = 0; i < NW; ++i) Wn.a[r][i] = __funnelshift_r(w[i], w[i + 1], sh);
    }
    return Wn;
}
/// two bytes (row ra, byte ba) and (row rb, byte bb) of the window into the low half of a word (one PRMT)
template<int CH>
__device__ __forceinline__ uint32_t win_pick2(const Window5<CH>& Wn, int ra, int ba, int rb, int bb) {
    return __byte_perm(Wn.a[ra][ba >> 2], Wn.a[rb][bb >> 2], (uint32_t)((ba & 3) | ((4 + (bb & 3)) << 4)));
}
/// the 16 LBSP neighbours of channel c in pattern order (LBSP.hpp:292-294); rows r = dy+2, bytes (dx+2)*CH + c
template<int CH>
__device__ __forceinline__ Lookup16 lbsp_lookup_window(const Window5<CH>& Wn, int c) {
#define LVB_B(dx) (((dx) + 2) * CH + c)
    Lookup16 L;
    L.w[0] = __byte_perm(win_pick2<CH>(Wn, 2, LVB_B(-2), 2, LVB_B(2)), win_pick2<CH>(Wn, 0, LVB_B(0), 4, LVB_B(0)), 0x5410);   // (-2,0) (2,0) (0,-2) (0,2)
    L.w[1] = __byte_perm(win_pick2<CH>(Wn, 4, LVB_B(-2), 0, LVB_B(2)), win_pick2<CH>(Wn, 4, LVB_B(2), 0, LVB_B(-2)), 0x5410);  // (-2,2) (2,-2) (2,2) (-2,-2)
    L.w[2] = __byte_perm(win_pick2<CH>(Wn, 3, LVB_B(0), 2, LVB_B(-1)), win_pick2<CH>(Wn, 1, LVB_B(0), 2, LVB_B(1)), 0x5410);   // (0,1) (-1,0) (0,-1) (1,0)
    L.w[3] = __byte_perm(win_pick2<CH>(Wn, 1, LVB_B(-1), 3, LVB_B(1)), win_pick2<CH>(Wn, 1, LVB_B(1), 3, LVB_B(-1)), 0x5410);  // (-1,-1) (1,1) (1,-1) (-1,1)
#undef LVB_B
    return L;
}
template<int CH>
__device__ __forceinline__ uint32_t win_center(const Window5<CH>& Wn, int c) {
    const int b = 2 * CH + c;
    return (Wn.a[2][b >> 2] >> (8 * (b & 3))) & 0xFFu;
}

/// desc = sum_n (|val_n - ref| > t) << n   (strict >, unsigned) — LBSP.hpp:193-224.
/// VABSDIFF4 + SWAR compare give one flag byte per neighbour; two dp4a chains with weights 1,2,4,..,128 gather
/// the 16 flags into the descriptor bits.
/// SMALL_T: every threshold is known to be <= 127 (the LUT cap sat(off + 255*rel) is 85 with the reference's defaults), so
/// |d| > t  <=>  bit 7 of ((d & 0x7f) + (127 - t))  OR  bit 7 of d : 3 instructions per 4 neighbours instead of 5, and the
/// flags stay at bit 7 (the dp4a sums are 128 x the descriptor halves, folded into the final shift).
template<bool SMALL_T = false>
__device__ __forceinline__ uint32_t lbsp_threshold(const Lookup16& L, uint32_t ref, uint32_t t) {
    const uint32_t r4 = __byte_perm(ref, 0, 0x0000);
    if constexpr (SMALL_T) {
        const uint32_t k4 = __byte_perm(127u - t, 0, 0x0000);
        uint32_t f[4];
#pragma unroll
        for(int q = 0; q < 4; ++q) {
            const uint32_t d = __vabsdiffu4(L.w[q], r4);
            f[q] = (((d & 0x7f7f7f7fu) + k4) | d) & 0x80808080u;
        }
        const uint32_t lo = __dp4a(f[1], 0x80402010u, __dp4a(f[0], 0x08040201u, 0u));
        const uint32_t hi = __dp4a(f[3], 0x80402010u, __dp4a(f[2], 0x08040201u, 0u));
        return (lo >> 7) | (hi << 1);
    } else {
        const uint32_t t4 = __byte_perm(t, 0, 0x0000);
        const uint32_t f0 = __vsetgtu4(__vabsdiffu4(L.w[0], r4), t4), f1 = __vsetgtu4(__vabsdiffu4(L.w[1], r4), t4);
        const uint32_t f2 = __vsetgtu4(__vabsdiffu4(L.w[2], r4), t4), f3 = __vsetgtu4(__vabsdiffu4(L.w[3], r4), t4);
        const uint32_t lo = __dp4a(f1, 0x80402010u, __dp4a(f0, 0x08040201u, 0u));
        const uint32_t hi = __dp4a(f3, 0x80402010u, __dp4a(f2, 0x08040201u, 0u));
        return lo | (hi << 8);
    }
}

/// x % n for x < 2^32 with magic = floor(2^32 / n): the estimated quotient is at most one too small
__device__ __forceinline__ uint32_t fast_mod(uint32_t x, uint32_t n, uint32_t magic) {
    uint32_t r = x - __umulhi(x, magic) * n;
    if(n == 1u) return 0u; // floor(2^32/1) does not fit: handled apart
    return r >= n ? r - n : r;
}

/// floor(x / n) with the same magic (exact: the estimate is corrected by at most one)
__device__ __forceinline__ uint32_t fast_div(uint32_t x, uint32_t n, uint32_t magic) {
    if(n == 1u) return x;
    const uint32_t q = __umulhi(x, magic);
    return (x - q * n) >= n ? q + 1u : q;
}

// ---------------------------------------------------------------------------------------------
// sampling patterns
// ---------------------------------------------------------------------------------------------
__device__ __constant__ const unsigned char c_pat7[49] = {
    2, 4, 6, 7, 6, 4, 2, 4, 8, 12, 14, 12, 8, 4, 6, 12, 21, 25, 21, 12, 6, 7, 14, 25, 28, 25, 14, 7,
    6, 12, 21, 25, 21, 12, 6, 4, 8, 12, 14, 12, 8, 4, 2, 4, 6, 7, 6, 4, 2};

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/// opencv.hpp:873-891,909-925 — walk the 7x7 integer Gaussian, clamp to [2,dim-3]
__device__ __forceinline__ void sample_pos_7x7(uint32_t rnd, int& sx, int& sy, int ox, int oy, int W, int H) {
    int r = 1 + (int)(rnd % 512u);
    int i = 0;
    for(; i < 49; ++i) { r -= c_pat7[i]; if(r <= 0) break; }
    if(i > 48) i = 48; // unreachable (weights sum to 512)
    sx = clampi(ox + (i % 7) - 3, 2, W - 3);
    sy = clampi(oy + (i / 7) - 3, 2, H - 3);
}
/// opencv.hpp:941-966 — offset of entry r of the 3x3 (8 entries) / 5x5 (24 entries) neighbour patterns, computed
/// arithmetically (row-major from +y, centre skipped) so divergent lanes do not serialise on a constant table
__device__ __forceinline__ void neighbor_offset(bool use3x3, uint32_t rnd, int& dx, int& dy) {
    if(use3x3) { const int r = (int)(rnd % 8u), k = r + (r >= 4); dx = k % 3 - 1; dy = 1 - k / 3; }
    else { const int r = (int)(rnd % 24u), k = r + (r >= 12); dx = k % 5 - 2; dy = 2 - k / 5; }
}

// ---------------------------------------------------------------------------------------------
// bit-packed rows (1 bit / pixel, 32 pixels / word, row pitch WW words, bits >= W are kept zero)
// ---------------------------------------------------------------------------------------------
enum { FILL_ZERO = 0, FILL_ONE = 1, FILL_REPL = 2 };

template<int FILL>
__device__ __forceinline__ uint32_t row_word(const uint32_t* __restrict__ row, int wi, int WW, int W) {
    if(wi < 0) {
        if(FILL == FILL_ZERO) return 0u;
        if(FILL == FILL_ONE) return 0xFFFFFFFFu;
        return (row[0] & 1u) ? 0xFFFFFFFFu : 0u;
    }
    const int rem = W & 31;
    if(wi >= WW) {
        if(FILL == FILL_ZERO) return 0u;
        if(FILL == FILL_ONE) return 0xFFFFFFFFu;
        const uint32_t last = row[WW - 1] >> ((W - 1) & 31);
        return (last & 1u) ? 0xFFFFFFFFu : 0u;
    }
    uint32_t w = row[wi];
    if(rem && wi == WW - 1) {
        const uint32_t valid = (1u << rem) - 1u;
        if(FILL == FILL_ONE) w |= ~valid;
        else if(FILL == FILL_REPL) { if((w >> (rem - 1)) & 1u) w |= ~valid; }
    }
    return w;
}
/// word whose bit i = source bit (i+d) of the (left,cur,right) triple, |d| <= 31
__device__ __forceinline__ uint32_t shift_bits(uint32_t left, uint32_t cur, uint32_t right, int d) {
    if(d == 0) return cur;
    return d > 0 ? __funnelshift_r(cur, right, d) : __funnelshift_l(left, cur, -d);
}
template<int R, bool DILATE>
__device__ __forceinline__ uint32_t hmorph(uint32_t left, uint32_t cur, uint32_t right) {
    uint32_t acc = cur;
#pragma unroll
    for(int d = 1; d <= R; ++d) {
        if(DILATE) acc |= shift_bits(left, cur, right, d) | shift_bits(left, cur, right, -d);
        else acc &= shift_bits(left, cur, right, d) & shift_bits(left, cur, right, -d);
    }
    return acc;
}

// ---------------------------------------------------------------------------------------------
// TMA + mbarrier (PTX ISA 8.x, sm_90+/sm_100a)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "LVB_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra LVB_DONE_%=;\n"
        "bra LVB_WAIT_%=;\n"
        "LVB_DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// ---- cp.async (LDGSTS): global -> shared copies that hold no registers while in flight (software-pipelined tile loops)
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
/// 4-byte copy; `valid` false writes zeros instead (the source address must still be a mapped one)
__device__ __forceinline__ void cp_async4_zfill(void* dst, const void* src, bool valid) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(valid ? 4 : 0) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template<int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

/// 2-D tiled bulk tensor load global -> shared, completion counted on `bar` (out-of-bounds elements read as 0)
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}

} // namespace lvb
