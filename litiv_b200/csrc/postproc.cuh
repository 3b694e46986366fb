// litiv_b200 — mask post-processing on bit-packed planes (1 bit / pixel), replacing the ~20 full-frame OpenCV
// calls of the reference (video/src/BackgroundSubtractorSuBSENSE.cpp:536-554, PAWCS.cpp:1443-1461,
// LOBSTER.cpp:578). Semantics per SURVEY.md Appendix E:
//   dilate/erode: rect kernel, pixels outside the image ignored; median on {0,255}: majority with replicated
//   border; floodFill((0,0)) + bitwise_not: background pixels not 4-connected to the image border ring.
#pragma once
#include "state.cuh"

namespace lvb {

struct PostArgs {
    int W, H, WW, Wp;
    uint32_t* raw; uint32_t* lastraw; uint32_t* lastrawblink; uint32_t* blinks;
    uint32_t* tmpA; uint32_t* pre; uint32_t* reach; uint32_t* comb;
    uint32_t* lastfg; uint32_t* dilinv; uint32_t* dil; // dil: optional copy of the dilated mask (PAWCS refreshModel reads it)
    uchar* out_mask; size_t out_pitch;
    float2* fin;               // final-segmentation EMAs (written)
    const float2* fin_in;      // ... read (== fin unless the caller ping-pongs the map)
    FrameCtl* ctl;
    int median_k;
    uint32_t frame; int avg_samples; // frame != 0: the EMA factors are derived from the frame index instead of being read from ctl
    // PAWCS only (null otherwise): the illumination mask of the next frame, new[p] = did[p+1] ? (roi[p]==255) : did[p] (snapshot semantics,
    // DESIGN.md section 2), is computed by pp_blink_close, which runs over the same word grid
    const uint32_t* did; const uint32_t* roi255; const uint32_t* roi; uint32_t* illum;
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

/// P1 (blink bookkeeping, SuBSENSE.cpp:536-539) + the whole MORPH_CLOSE (:540) in one launch: erode3x3(dilate3x3(raw)) needs raw rows y-2..y+2; the dilated rows are
/// recomputed per word (a handful of bit operations each) instead of round-tripping through a plane and a launch boundary.
__global__ void __launch_bounds__(256) pp_blink_close(const PostArgs A) {
    const int wi = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if(wi >= A.WW) return;
    const size_t i = (size_t)y * A.WW + wi;
    const uint32_t raw = A.raw[i];
    const uint32_t cur_blink = raw ^ A.lastraw[i];
    A.blinks[i] = cur_blink | A.lastrawblink[i];
    A.lastrawblink[i] = cur_blink;
    A.lastraw[i] = raw;
    // erode: AND over rows y-1..y+1 (rows outside the image are ignored) of the horizontally eroded dilated row; the
    // horizontal erosion needs the dilated words wi-1, wi, wi+1 (outside the image: all ones)
    uint32_t acc = 0xFFFFFFFFu;
#pragma unroll
    for(int dy = -1; dy <= 1; ++dy) {
        const int yy = y + dy;
        if(yy < 0 || yy >= A.H) continue;
        uint32_t d[3];
#pragma unroll
        for(int k = -1; k <= 1; ++k) {
            const int w2 = wi + k;
            d[k + 1] = (w2 < 0 || w2 >= A.WW) ? 0xFFFFFFFFu : morph_word<1, true>(A.raw, yy, w2, A.H, A.WW, A.W);
        }
        if(wi == A.WW - 1 && (A.W & 31)) d[1] |= ~valid_mask(wi, A.WW, A.W); // bits beyond the row end count as "outside"
        if(wi + 1 == A.WW - 1 && (A.W & 31)) d[2] |= ~valid_mask(wi + 1, A.WW, A.W);
        acc &= hmorph<1, false>(d[0], d[1], d[2]);
    }
    A.pre[i] = acc & valid_mask(wi, A.WW, A.W);
    if(A.illum) {
        const uint32_t d = A.did[i], nxt = wi + 1 < A.WW ? A.did[i + 1] : 0u;
        const uint32_t dn = (d >> 1) | (nxt << 31); // bit x = did[x+1]
        A.illum[i] = ((dn & A.roi255[i]) | (~dn & d)) & A.roi[i];
    }
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

// ---- hole filling: connected components of the background on ROW RUNS, lock-free union-find --------------------------
// cv::floodFill(pre,(0,0),255) + bitwise_not marks the background pixels that are NOT 4-connected to the image border
// (the 2-px border ring is always background). A background run of row y and one of row y-1 are connected iff they
// share a column, i.e. each maximal run of (bg[y-1] & bg[y]) is one edge. Nodes: 0 = border, 1 + y*RS + rank = the
// rank-th background run of row y. Three tiny kernels (warp per row) replace the iterative flood; cost is independent
// of the blob geometry.
struct HoleArgs {
    int W, H, WW, RS;
    const uint32_t* pre; const uint32_t* raw; uint32_t* comb;
    uint32_t* parent; ushort* rankbase;   // parent[1 + H*RS], rankbase[H*WW]
};
__device__ __forceinline__ uint32_t bg_word(const uint32_t* __restrict__ pre, int y, int wi, int WW, int W) {
    return (wi < 0 || wi >= WW) ? 0u : (~pre[(size_t)y * WW + wi] & valid_mask(wi, WW, W));
}
/// start bits of the background runs inside word wi (a run continuing from the previous word has no start bit here)
__device__ __forceinline__ uint32_t run_starts(uint32_t m, uint32_t m_prev) { return m & ~((m << 1) | (m_prev >> 31)); }

__device__ __forceinline__ uint32_t uf_find(uint32_t* P, uint32_t x) {
    uint32_t p = __ldcg(P + x);
    while(p != x) {
        const uint32_t gp = __ldcg(P + p);
        if(gp != p) __stcg(P + x, gp); // path halving: only ever re-points to an ancestor
        x = p; p = gp;
    }
    return x;
}
__device__ __forceinline__ void uf_union(uint32_t* P, uint32_t a, uint32_t b) {
    while(true) {
        a = uf_find(P, a); b = uf_find(P, b);
        if(a == b) return;
        if(a < b) { const uint32_t t = a; a = b; b = t; } // hook the larger root under the smaller: node 0 (border) stays a root
        if(atomicCAS(P + a, a, b) == a) return;
    }
}

/// UF1: per row, number the background runs and make each its own root -- except the runs that touch the image border (every
/// run of the first / last row, the run holding the first pixel, the run holding the last pixel): they start directly under node 0
/// (parent[0] == 0 from the allocation; node 0 is never hooked under anything). In a sparse mask most runs span their row, so the
/// vertical unions of UF2 then find equal roots at once instead of building row-to-row chains up to H long.
__global__ void __launch_bounds__(256) pp_holes_init(const HoleArgs A) {
    const int y = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if(y >= A.H) return;
    const bool all_touch = (y == 0 || y == A.H - 1);
    const int nchunks = (A.WW + 31) >> 5;
    uint32_t base = 0;
    bool first_bg = false, last_bg = false;
    for(int k = 0; k < nchunks; ++k) {
        const int wi = k * 32 + lane;
        const uint32_t m = bg_word(A.pre, y, wi, A.WW, A.W);
        const uint32_t st = run_starts(m, bg_word(A.pre, y, wi - 1, A.WW, A.W));
        uint32_t cnt = __popc(st), incl = cnt;
#pragma unroll
        for(int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, o); if(lane >= o) incl += v; }
        const uint32_t excl = base + incl - cnt;
        if(wi < A.WW) {
            A.rankbase[(size_t)y * A.WW + wi] = (ushort)excl;
            for(uint32_t r = 0; r < cnt; ++r) { const uint32_t id = 1u + (uint32_t)y * A.RS + excl + r; A.parent[id] = all_touch ? 0u : id; }
        }
        first_bg |= __any_sync(0xFFFFFFFFu, wi == 0 && (m & 1u));
        last_bg |= __any_sync(0xFFFFFFFFu, wi == A.WW - 1 && ((m >> ((A.W - 1) & 31)) & 1u));
        base += __shfl_sync(0xFFFFFFFFu, incl, 31);
    }
    __syncwarp();
    if(lane == 0 && !all_touch && base > 0u) {
        if(first_bg) A.parent[1u + (uint32_t)y * A.RS] = 0u;                 // the run that starts at x = 0 is the row's first
        if(last_bg) A.parent[1u + (uint32_t)y * A.RS + base - 1u] = 0u;      // the run that reaches x = W-1 is the row's last
    }
}
/// node id of the background run of row y that contains bit b of word wi
__device__ __forceinline__ uint32_t run_id(const HoleArgs& A, int y, int wi, int b) {
    const uint32_t st = run_starts(bg_word(A.pre, y, wi, A.WW, A.W), bg_word(A.pre, y, wi - 1, A.WW, A.W));
    const uint32_t upto = b == 31 ? 0xFFFFFFFFu : ((2u << b) - 1u);
    return 1u + (uint32_t)y * A.RS + A.rankbase[(size_t)y * A.WW + wi] + __popc(st & upto) - 1u;
}
/// UF2: union vertically adjacent runs (the border contacts were settled by UF1)
__global__ void __launch_bounds__(256) pp_holes_union(const HoleArgs A) {
    const int y = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if(y >= A.H) return;
    if(y > 0) {
        for(int wi = lane; wi < A.WW; wi += 32) {
            const uint32_t m = bg_word(A.pre, y, wi, A.WW, A.W);
            const uint32_t c = m & bg_word(A.pre, y - 1, wi, A.WW, A.W);
            const uint32_t cp = bg_word(A.pre, y, wi - 1, A.WW, A.W) & bg_word(A.pre, y - 1, wi - 1, A.WW, A.W);
            uint32_t cs = run_starts(c, cp);
            while(cs) { const int b = __ffs(cs) - 1; cs &= cs - 1; uf_union(A.parent, run_id(A, y - 1, wi, b), run_id(A, y, wi, b)); }
        }
    }
}
/// UF3: runs whose root is the border node are "reached"; fill them from their start bits (carry trick), the rest of the
/// background is holes; fused with  m = raw | holes | erode7x7(pre)  (SuBSENSE.cpp:543-546)
__global__ void __launch_bounds__(256) pp_holes_combine(const HoleArgs A) {
    const int y = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if(y >= A.H) return;
    const int nchunks = (A.WW + 31) >> 5;
    uint32_t carry = 0;
    for(int k = 0; k < nchunks; ++k) {
        const int wi = k * 32 + lane;
        const uint32_t m = bg_word(A.pre, y, wi, A.WW, A.W);
        uint32_t st = run_starts(m, bg_word(A.pre, y, wi - 1, A.WW, A.W)), rs = 0;
        if(wi < A.WW) {
            const uint32_t base = 1u + (uint32_t)y * A.RS + A.rankbase[(size_t)y * A.WW + wi];
            uint32_t r = 0;
            while(st) { const int b = __ffs(st) - 1; st &= st - 1; if(uf_find(A.parent, base + r) == 0u) rs |= 1u << b; ++r; }
        }
        const uint32_t reach = fill_up_chunk(m, rs, carry);
        if(wi < A.WW) {
            const size_t i = (size_t)y * A.WW + wi;
            A.comb[i] = A.raw[i] | (m & ~reach) | morph_word<3, false>(A.pre, y, wi, A.H, A.WW, A.W);
        }
    }
}

/// holes only (standalone operator / tests): same as pp_holes_combine without the raw / erode terms
__global__ void __launch_bounds__(256) pp_holes_only(const HoleArgs A) {
    const int y = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if(y >= A.H) return;
    const int nchunks = (A.WW + 31) >> 5;
    uint32_t carry = 0;
    for(int k = 0; k < nchunks; ++k) {
        const int wi = k * 32 + lane;
        const uint32_t m = bg_word(A.pre, y, wi, A.WW, A.W);
        uint32_t st = run_starts(m, bg_word(A.pre, y, wi - 1, A.WW, A.W)), rs = 0;
        if(wi < A.WW) {
            const uint32_t base = 1u + (uint32_t)y * A.RS + A.rankbase[(size_t)y * A.WW + wi];
            uint32_t r = 0;
            while(st) { const int b = __ffs(st) - 1; st &= st - 1; if(uf_find(A.parent, base + r) == 0u) rs |= 1u << b; ++r; }
        }
        const uint32_t reach = fill_up_chunk(m, rs, carry);
        if(wi < A.WW) A.comb[(size_t)y * A.WW + wi] = m & ~reach;
    }
}
/// standalone rect dilate / erode (radius 1 or 3)
__global__ void __launch_bounds__(256) mask_morph_kernel(const uint32_t* __restrict__ src, uint32_t* dst, int W, int H, int WW, bool dilate, int r) {
    const int wi = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if(wi >= WW) return;
    uint32_t v;
    if(dilate) v = r == 1 ? morph_word<1, true>(src, y, wi, H, WW, W) : morph_word<3, true>(src, y, wi, H, WW, W);
    else v = r == 1 ? morph_word<1, false>(src, y, wi, H, WW, W) : morph_word<3, false>(src, y, wi, H, WW, W);
    dst[(size_t)y * WW + wi] = v;
}

/// k x k majority (== cv::medianBlur on a binary mask, replicated borders). Each thread owns one column of a
/// 32 x (8*MEDIAN_ROWS) tile and slides the window down it: the k-bit row popcounts enter and leave a running sum,
/// so a pixel costs ~2 row evaluations instead of k. Writes the bit-packed result and the caller's byte mask.
constexpr int MEDIAN_ROWS = 16, MEDIAN_MAXR = 15, MEDIAN_STAGE = MEDIAN_ROWS + 2 * MEDIAN_MAXR; // k <= 31
/// A warp owns one mask word (32 columns) x MEDIAN_ROWS rows. All the row words its windows can touch (rows y0-r .. y0+15+r,
/// words wi-1..wi+1, with the replicated-border rules applied) are staged in shared memory by ONE round of loads, so the sliding
/// sum runs without a dependent memory access per row; a lane's k-wide window is then one funnel shift away.
__global__ void __launch_bounds__(256) pp_median(const uint32_t* __restrict__ src, uint32_t* dst, uchar* out_mask, size_t out_pitch,
                                                  int W, int H, int WW, int k) {
    __shared__ uint32_t s_rows[8][MEDIAN_STAGE][3];
    const int lane = threadIdx.x;
    const int x = blockIdx.x * 32 + lane, y0 = (blockIdx.y * 8 + threadIdx.y) * MEDIAN_ROWS;
    const int wi = x >> 5, xb = x & 31, r = k >> 1, half = (k * k) / 2;
    const uint32_t wmask = k >= 32 ? 0xFFFFFFFFu : (1u << k) - 1u;
    const bool col_ok = x < W && wi < WW;
    if(y0 >= H || wi >= WW) return; // warp-uniform: y0 and wi depend on threadIdx.y / blockIdx only
    uint32_t (*rows)[3] = s_rows[threadIdx.y];
    const int nstage = MEDIAN_ROWS + 2 * r;  // staged row j <-> image row y0 - r + j (clamped: replicated border)
    for(int t = lane; t < nstage * 3; t += 32) {
        const int j = t / 3, w = t - j * 3;
        rows[j][w] = row_word<FILL_REPL>(src + (size_t)clampi(y0 - r + j, 0, H - 1) * WW, wi - 1 + w, WW, W);
    }
    __syncwarp();
    auto count = [&](int j) { // set pixels in this lane's window on staged row j
        const uint32_t l = rows[j][0], c = rows[j][1], rr = rows[j][2];
        const uint32_t win = (xb >= r) ? __funnelshift_r(c, rr, xb - r) : __funnelshift_r(l, c, 32 + xb - r);
        return __popc(win & wmask);
    };
    int sum = 0;
    for(int j = 0; j <= 2 * r; ++j) sum += count(j);
#pragma unroll 4
    for(int i = 0; i < MEDIAN_ROWS; ++i) {
        const int y = y0 + i;
        if(y >= H) break; // warp-uniform
        if(i > 0) sum += count(i + 2 * r) - count(i - 1);
        const bool on = col_ok && sum > half;
        if(col_ok && out_mask) out_mask[(size_t)y * out_pitch + x] = on ? 255 : 0;
        const uint32_t b = __ballot_sync(0xFFFFFFFFu, on);
        if(lane == 0) dst[(size_t)y * WW + wi] = b;
    }
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
    if(A.dil) A.dil[i] = dil;
}

/// final-segmentation EMAs (SuBSENSE.cpp:553-554): cv::addWeighted accumulates in double and rounds once
__global__ void __launch_bounds__(256) pp_final_ema(const PostArgs A) {
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if(x >= A.W || y >= A.H) return;
    // the frame tail advances ctl->aLT/aST; a caller that runs this kernel beside the tail passes the frame index instead
    const float aLT = A.frame ? __fdiv_rn(1.0f, (float)min(A.frame, (uint32_t)A.avg_samples)) : A.ctl->aLT;
    const float aST = A.frame ? __fdiv_rn(1.0f, (float)min(A.frame, (uint32_t)A.avg_samples / 4u)) : A.ctl->aST;
    const double v = ((A.lastfg[(size_t)y * A.WW + (x >> 5)] >> (x & 31)) & 1u) ? 255.0 : 0.0;
    const size_t pix = (size_t)y * A.Wp + x;
    float2 f = A.fin_in[pix];
    f.x = (float)__dadd_rn(__dmul_rn((double)f.x, (double)__fsub_rn(1.0f, aLT)), __dmul_rn(v, __dmul_rn(1.0 / 255, (double)aLT)));
    f.y = (float)__dadd_rn(__dmul_rn((double)f.y, (double)__fsub_rn(1.0f, aST)), __dmul_rn(v, __dmul_rn(1.0 / 255, (double)aST)));
    A.fin[pix] = f;
}

} // namespace lvb
