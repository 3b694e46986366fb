// litiv_b200 — SuBSENSE per-frame kernels (replaces BackgroundSubtractorSuBSENSE_::apply,
// reference video/src/BackgroundSubtractorSuBSENSE.cpp:188-612, and refreshModel :80-105).
//
// Frame = phase A (classification + feedback, 1 thread / pixel, reads only frame-start state)
//       -> phase B (deferred neighbour-sample writes, gathered per target pixel in raster order of the source)
//       -> bit-packed post-processing (postproc.cuh) -> frame tail (LUT adaptation, frame-level reset logic)
//       -> conditional refresh.
#pragma once
#include "state.cuh"

namespace lvb {

#ifndef LVB_TILE_H
#define LVB_TILE_H 8
#endif
constexpr int TILE_W = 32, TILE_H = LVB_TILE_H, HALO = 2; // tile of the TMA-staged kernels (one warp per tile row)
#ifndef PHASEA_MIN_BLOCKS
#define PHASEA_MIN_BLOCKS 4
#endif
constexpr int TILE_ROWS = TILE_H + 2 * HALO;
#ifndef LVB_SCAN_GRID_CTAS
#define LVB_SCAN_GRID_CTAS 16
#endif
constexpr int SCAN_GRID_CTAS_PER_SM = LVB_SCAN_GRID_CTAS; // grid of the multi-tile scan kernel, in CTAs per SM (4 are resident)
// Programmatic dependent launch between the scan kernel and the two tail passes (-DLVB_PDL=0 compiles it out): a tail pass is launched while its
// predecessor drains, stages its LUT, and waits in pdl_wait() for the predecessor's memory before it touches the work-list
// (1080p RGB: frame 0.2597 -> 0.2563 ms, end to end 7.16 -> 7.35 Gpx/s; LVB_NO_PDL=1 launches them the plain way).
#ifndef LVB_PDL
#define LVB_PDL 1
#endif
__device__ __forceinline__ void pdl_wait() {
#if LVB_PDL
    asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}
__device__ __forceinline__ void pdl_launch_dependents() {
#if LVB_PDL
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
constexpr uint32_t NO_INTENT = 0xFFFFu; // intents[] entry of a pixel that queued no neighbour write (valid codes are <= 0x18FF)
// TMA boxes must start on a 16-byte boundary of the image row: the box starts TILE_SHIFT bytes before the halo's first
// byte ((32k-2)*ch mod 16 is the same for every tile) and is TILE_SHIFT bytes wider.
__host__ __device__ constexpr int tile_shift(int ch) { return (16 - (HALO * ch) % 16) % 16; }
__host__ __device__ constexpr int tile_pitch(int ch) { return (tile_shift(ch) + (TILE_W + 2 * HALO) * ch + 15) / 16 * 16; }

/// stage the (TILE_W+4)x(TILE_H+4) input tile into shared memory: one TMA bulk tensor copy (zero-filled
/// outside the image) or, when the frame pitch is not TMA-compatible, a cooperative byte copy.
/// begin() issues the copy, wait() blocks until the tile is visible to the whole CTA; independent global loads
/// issued between the two overlap with the TMA transfer.
template<int CH>
__device__ __forceinline__ void stage_tile_begin(uchar* tile, uint64_t* bar, const CUtensorMap* tmap, int use_tma,
                                                 const uchar* img, size_t ipitch, int W, int H, int x0, int y0) {
    constexpr int PITCH = tile_pitch(CH);
    if(use_tma) {
        if(threadIdx.x == 0 && threadIdx.y == 0) {
            mbar_init(bar, 1);
            mbar_fence_init();
            mbar_expect_tx(bar, PITCH * TILE_ROWS);
            tma_load_2d(tile, tmap, (x0 - HALO) * CH - tile_shift(CH), y0 - HALO, bar);
        }
    } else {
        const int tid = threadIdx.y * blockDim.x + threadIdx.x, nt = blockDim.x * blockDim.y;
        const int rowbytes = (TILE_W + 2 * HALO) * CH;
        for(int i = tid; i < rowbytes * TILE_ROWS; i += nt) {
            const int r = i / rowbytes, b = i - r * rowbytes;
            const int gy = y0 - HALO + r, gb = (x0 - HALO) * CH + b;
            uchar v = 0;
            if(gy >= 0 && gy < H && gb >= 0 && gb < W * CH) v = img[(size_t)gy * ipitch + gb];
            tile[r * PITCH + tile_shift(CH) + b] = v;
        }
    }
}
__device__ __forceinline__ void stage_tile_wait(uint64_t* bar, int use_tma) {
    __syncthreads();               // mbarrier init (thread 0) / cooperative copy visible to everyone
    if(use_tma) mbar_wait(bar, 0); // TMA bytes have landed (async proxy -> visible after the phase flips)
}
template<int CH>
__device__ __forceinline__ void stage_tile(uchar* tile, uint64_t* bar, const CUtensorMap* tmap, int use_tma,
                                           const uchar* img, size_t ipitch, int W, int H, int x0, int y0) {
    stage_tile_begin<CH>(tile, bar, tmap, use_tma, img, ipitch, W, H, x0, y0);
    stage_tile_wait(bar, use_tma);
}

// ---- sample-consensus scan (SuBSENSE.cpp:229-253 / :367-395) ----
/// one sample against one pixel: colour gate, then the descriptor test. Returns whether the sample matches and
/// its total descriptor / colour+descriptor distances.
template<int CH, bool T7>
__device__ __forceinline__ bool subsense_test_sample(const Lookup16 (&L)[CH], const uint32_t (&cur)[CH], const uint32_t (&intra)[CH],
                                                     const typename Pack<CH>::Col bc, const typename Pack<CH>::Desc bd,
                                                     uint32_t thrC, uint32_t thrD, const uchar* s_lut, uint32_t& totDesc, uint32_t& totSum) {
    const uint32_t totC = thrC * 3u, totD = thrD * 3u, scC = totC >> 1;
    bool ok = true;
    uint32_t cd[CH];
#pragma unroll
    for(int c = 0; c < CH; ++c) {
        const uint32_t b = col_get(bc, c);
        cd[c] = cur[c] > b ? cur[c] - b : b - cur[c];
        ok = ok && (cd[c] <= (CH == 1 ? thrC : scC));
    }
    totDesc = 0; totSum = 0;
    if(ok) {
#pragma unroll
        for(int c = 0; c < CH; ++c) {
            const uint32_t b = col_get(bc, c), d = desc_get(bd, c);
            const uint32_t inter = lbsp_threshold<T7>(L[c], b, s_lut[b]);
            const uint32_t dd = (__popc(intra[c] ^ d) + __popc(inter ^ d)) >> 1;
            if(CH == 1) {
                const uint32_t sum = min((dd >> 2) * 15u + cd[c], 255u);
                ok = ok && (dd <= thrD) && (sum <= thrC);
                totDesc = dd; totSum = sum;
            } else {
                const uint32_t sum = min((dd >> 1) * 15u + cd[c], 255u);
                ok = ok && (sum <= scC);
                totDesc += dd; totSum += sum;
            }
        }
        if(CH != 1) ok = ok && !(totDesc > totD || totSum > totC);
    }
    return ok;
}

/// Work-list entry of a pixel that is still undecided after the two prefetched samples. The scan kernel appends it (warp-aggregated
/// atomic on FrameCtl::wl_count) and moves on: the rest of the pixel's scan runs in subsense_tail_pass, spread over the whole chip,
/// instead of holding the 32x8 tile's CTA through a chain of dependent DRAM round trips. Only ~17 % of the pixels get here (3 or more
/// samples), ~2-6 % are foreground and scan all N. Structure of arrays [FIELDS][wl_cap] u32: consecutive entries are written by
/// consecutive lanes of one warp.
template<int CH> struct WlCtx {
    static constexpr int LOOK = 0;                      // 4*CH words: the 16 LBSP neighbours of every channel
    static constexpr int CUR = 4 * CH;                  // packed colour
    static constexpr int INTRA = CUR + 1;               // intra descriptors (CH==3: 2 words)
    static constexpr int THRC = INTRA + (CH == 3 ? 2 : 1), THRD = THRC + 1, PIX = THRD + 1;
    static constexpr int STATE = PIX + 1;               // minSum (10 bits) | minDesc << 10 (6) | good << 16 (8) | scanned << 24 (8)
    static constexpr int FIELDS = (STATE + 1 + 3) / 4 * 4;   // array of structures, entries padded to 16 bytes: a warp's appends are one contiguous run of full sectors
    static constexpr int VEC = FIELDS / 4;
};
__device__ __forceinline__ uint32_t wl_state_pack(uint32_t minSum, uint32_t minDesc, uint32_t good, uint32_t s) { return minSum | (minDesc << 10) | (good << 16) | (s << 24); }

/// slice of FrameCtl staged in shared memory by the prologue (the feedback step reads it long after the state loads)
struct CtlSlice { float aLT, aST, t_lower, t_upper; uint32_t frame, cooldown, use3x3, pad; };
constexpr int GHOST_ROWS = TILE_H + 2 * HALO;

// The reference's per-pixel loop (SuBSENSE.cpp:201-481) is split in two kernels so that the mask post-processing chain of
// frame k (which only needs the raw mask) runs on a side stream beside the feedback step of frame k and the scan of frame
// k+1, instead of sitting between two frames:
//   scan     (A1): LBSP + sample-consensus scan -> raw mask, per-pixel hand-off word, last colour / descriptor
//   feedback (A2): D_min / T / v / R maps, stochastic own-sample write, queued neighbour write, frame tail (last CTA)
// Hand-off (uint2 per pixel): x = minSum | minDesc << 16 ; y = good | lastL1 << 16 | lastHd << 24.

/// Tensor maps of the per-pixel planes the scan kernel stages with TMA besides the input tile (built once per instance; the
/// colour / descriptor planes are ping-ponged between frames, so there is a map per plane and the host passes the pair in use)
struct ScanMaps { CUtensorMap pcol, pdesc, intents, own, rpl, bg; };
/// shared-memory layout of one scan tile: every operand is its own TMA box (128-byte aligned start, rows of a 16-byte multiple)
template<int CH> struct ScanTile {
    typedef typename Pack<CH>::Col Col;
    typedef typename Pack<CH>::Desc Desc;
    typedef typename Pack<CH>::Rec Rec;
    static constexpr int IW = TILE_W + 2 * HALO, IH = TILE_H + 2 * HALO;                  // 36 x 12: tile + 2-px halo
    // a TMA box has to START on a 16-byte boundary of the plane row (and span a 16-byte multiple): the halo'd boxes begin LP_x >= HALO
    // elements left of the tile (x0 is a multiple of 32) and are BW_x elements wide; pixel x0 - HALO + k sits at column k + LP_x - HALO
    static constexpr int lpad(int esz) { return esz >= 16 ? HALO : (HALO + 16 / esz - 1) / (16 / esz) * (16 / esz); }
    static constexpr int bwid(int esz) { return esz >= 16 ? IW : (lpad(esz) + TILE_W + HALO + 16 / esz - 1) / (16 / esz) * (16 / esz); }
    static constexpr int LP_COL = lpad(sizeof(Col)), LP_DESC = lpad(sizeof(Desc)), LP_INT = lpad(2);
    static constexpr int BW_COL = bwid(sizeof(Col)), BW_DESC = bwid(sizeof(Desc)), BW_INT = bwid(2);
    static constexpr int SH_COL = LP_COL - HALO, SH_DESC = LP_DESC - HALO, SH_INT = LP_INT - HALO;   // column shift of halo index k
    static constexpr int BG_ELEM = sizeof(Rec) >= 8 ? 8 : 4;                              // the sample planes are described as u64 / u32 elements (a TMA box row holds <= 256 elements)
    static constexpr int a128(int v) { return (v + 127) / 128 * 128; }
    static constexpr int O_IMG = 0;
    static constexpr int O_PCOL = a128(O_IMG + tile_pitch(CH) * TILE_ROWS);
    static constexpr int O_PDESC = a128(O_PCOL + BW_COL * IH * (int)sizeof(Col));
    static constexpr int O_INT = a128(O_PDESC + BW_DESC * IH * (int)sizeof(Desc));
    static constexpr int O_OWN = a128(O_INT + BW_INT * IH * 2);
    static constexpr int O_R = a128(O_OWN + TILE_W * TILE_H);
    static constexpr int O_SMP = a128(O_R + TILE_W * TILE_H * 4);
    static constexpr int BYTES = O_SMP + 2 * TILE_W * TILE_H * (int)sizeof(Rec);
    static constexpr uint32_t TX_PLANES = (uint32_t)(BW_COL * IH * sizeof(Col) + BW_DESC * IH * sizeof(Desc) + TILE_W * TILE_H * 4 + 2 * TILE_W * TILE_H * sizeof(Rec));
    static constexpr uint32_t TX_PENDING = (uint32_t)(BW_INT * IH * 2 + TILE_W * TILE_H);
};

// ---- colour bounding box of a pixel's samples (SuBSENSE scan: a pure acceleration structure, results unchanged) ------------------------------
// x = per-channel minimum, y = per-channel maximum of the colours of the pixel's N samples (bytes 0..2; 1 channel: byte 0), CONSERVATIVE: it
// contains every sample (it may be larger: it only grows between rebuilds). A sample can only match when every channel of the current colour is
// within the single-channel colour gate of the sample (SuBSENSE.cpp:232-235 / :371-377) and the summed distance within the total gate, so when
// the current colour lies farther from the box than that, NO sample matches: the pixel is foreground with zero matches and default minimal
// distances, exactly what scanning all N samples yields, without reading them (94-96 % of the foreground pixels of the SURVEY 8(d) sequence).
// Kept in step with every sample write: queued own / neighbour writes (scan kernel, standalone phase B, refresh kernel) grow it, a refresh
// rebuilds it from the samples it leaves behind, state import and a periodic pass (every 128 frames) rebuild it exactly (cbox_rebuild_kernel).
__device__ __forceinline__ uint2 cbox_full() { return make_uint2(0u, 0x00FFFFFFu); }       // contains everything: never filters
__device__ __forceinline__ uint2 cbox_empty() { return make_uint2(0x00FFFFFFu, 0u); }
/// colour inside the box? per byte |c - mn| + |mx - c| == mx - mn exactly when mn <= c <= mx (two native VABSDIFF4; a byte that lies outside
/// exceeds its width by an even amount, so a carry from the byte below cannot make it look equal; an empty box is never "contained")
__device__ __forceinline__ bool cbox_contains(const uint2 b, uint32_t col32) {
    col32 &= 0x00FFFFFFu;
    return __vabsdiffu4(col32, b.x) + __vabsdiffu4(b.y, col32) == b.y - b.x;
}
/// (the per-byte min / max / saturating-subtract intrinsics are emulated on this architecture: they sit behind the containment test,
/// which is what almost every colour written to a model passes)
__device__ __forceinline__ void cbox_grow(uint2& b, uint32_t col32) {
    if(cbox_contains(b, col32)) return;
    col32 &= 0x00FFFFFFu; b.x = __vminu4(b.x, col32); b.y = __vmaxu4(b.y, col32);
}
/// true when no colour inside the box can pass the colour gates against `cur32` (gate: per channel, tot: sum over the channels)
template<int CH>
__device__ __forceinline__ bool cbox_excludes(const uint2 b, uint32_t cur32, uint32_t gate, uint32_t tot) {
    if(cbox_contains(b, cur32)) return false;   // (most background pixels)
    cur32 &= 0x00FFFFFFu;
    const uint32_t d4 = __vmaxu4(__vsubus4(b.x, cur32), __vsubus4(cur32, b.y)); // per channel: distance from the box (0 inside)
    const uint32_t g4 = min(gate, 255u) * 0x00010101u;
    if(__vcmpgtu4(d4, g4) & 0x00FFFFFFu) return true;
    return CH != 1 && __dp4a(d4, 0x00010101u, 0u) > tot;
}
/// exact box of the N samples of one pixel
template<int CH>
__device__ __forceinline__ uint2 cbox_of_model(const void* bg, size_t plane, size_t pix, int N) {
    typedef typename Pack<CH>::Rec Rec;
    uint2 b = cbox_empty();
    for(int s = 0; s < N; ++s) cbox_grow(b, col_as_u32_(rec_col(((const Rec*)bg)[(size_t)s * plane + pix])));
    return b;
}
template<int CH>
__global__ void __launch_bounds__(256) cbox_rebuild_kernel(const void* bg, size_t plane, int W, int H, int Wp, int N, uint2* cbox) {
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if(x >= W || y >= H) return;
    const size_t pix = (size_t)y * Wp + x;
    cbox[pix] = cbox_of_model<CH>(bg, plane, pix, N);
}

/// Scan kernel: one thread per pixel of a 32x8 tile. EVERYTHING the tile reads (input tile + halo, previous colour / descriptor
/// tiles + halo, R(x), the first two sample records of every pixel, and - when neighbour / own-sample writes of the previous frame are
/// pending - the intent tile + halo and the own-slot tile) is staged by one elected thread with cp.async.bulk.tensor (TMA) behind a
/// single mbarrier: no thread issues a global load in the prologue, and the queued sample writes find their source records in shared memory.
template<int CH, bool T7>
__global__ void __launch_bounds__(TILE_W * TILE_H, PHASEA_MIN_BLOCKS)
subsense_scan(const SubArgs A, const __grid_constant__ CUtensorMap tmap, const __grid_constant__ ScanMaps M) {
    typedef typename Pack<CH>::Col Col;
    typedef typename Pack<CH>::Desc Desc;
    typedef typename Pack<CH>::Rec Rec;
    typedef WlCtx<CH> X;
    typedef ScanTile<CH> ST;
    constexpr int PITCH = tile_pitch(CH);
    constexpr int IW = ST::IW, IH = ST::IH;
    // PERSISTENT: a CTA walks tiles blockIdx.x, blockIdx.x + gridDim.x, ... ; the boxes of tile i+1 are in flight (second buffer, second
    // mbarrier) while tile i is computed, so only the first tile of a CTA waits for DRAM
    __shared__ __align__(128) uchar s_bufs[2][ST::BYTES];
    __shared__ __align__(8) uint64_t s_bar[2];
    __shared__ uchar s_lut[256];
    __shared__ uint32_t s_cnt[5];                 // nonzero | scanned | - | fg | warps done
    // per target pixel: mask of the 5x5 window positions whose source aims a queued neighbour write at it (bit = (dy+2)*5 + k = 24 - offset index)
    __shared__ uint32_t s_hits2[2][TILE_H][TILE_W];

    const int tid = threadIdx.y * TILE_W + threadIdx.x;
    pdl_launch_dependents();
    const int tiles_x = A.Wp / TILE_W, ntiles = tiles_x * ((A.H + TILE_H - 1) / TILE_H);
    const int step_y = (int)gridDim.x / tiles_x, step_x = (int)gridDim.x - step_y * tiles_x;
    auto issue = [&](int tx, int ty, int b) { // thread 0 only
        uchar* buf = s_bufs[b];
        const int x0 = tx * TILE_W, y0 = ty * TILE_H;
        mbar_expect_tx(&s_bar[b], (A.use_tma ? (uint32_t)(PITCH * TILE_ROWS) : 0u) + ST::TX_PLANES + ST::TX_PENDING);
        if(A.use_tma) tma_load_2d(buf + ST::O_IMG, &tmap, (x0 - HALO) * CH - tile_shift(CH), y0 - HALO, &s_bar[b]);
        tma_load_2d(buf + ST::O_SMP, &M.bg, x0 * (int)(sizeof(Rec) / ST::BG_ELEM), y0, &s_bar[b]);                                  // sample 0
        tma_load_2d(buf + ST::O_SMP + TILE_W * TILE_H * sizeof(Rec), &M.bg, x0 * (int)(sizeof(Rec) / ST::BG_ELEM), A.H + y0, &s_bar[b]);   // sample 1 (plane rows H..2H-1)
        tma_load_2d(buf + ST::O_PCOL, &M.pcol, x0 - ST::LP_COL, y0 - HALO, &s_bar[b]);
        tma_load_2d(buf + ST::O_PDESC, &M.pdesc, x0 - ST::LP_DESC, y0 - HALO, &s_bar[b]);
        tma_load_2d(buf + ST::O_R, &M.rpl, x0, y0, &s_bar[b]);
        tma_load_2d(buf + ST::O_INT, &M.intents, x0 - ST::LP_INT, y0 - HALO, &s_bar[b]);   // (staged even when nothing is pending: 1.2 KB, and the
        tma_load_2d(buf + ST::O_OWN, &M.own, x0, y0, &s_bar[b]);                            //  `pending` flag's own load stays off the critical path)
    };
    int tile = blockIdx.x;
    int tx = tile % tiles_x, ty = tile / tiles_x;
    if(tid == 0) {
        mbar_init(&s_bar[0], 1); mbar_init(&s_bar[1], 1);
        mbar_fence_init();
        if(tile < ntiles) issue(tx, ty, 0);
    }
    for(int i = tid; i < 256; i += TILE_W * TILE_H) s_lut[i] = A.lut[i];
    if(tid < 5) s_cnt[tid] = 0;
    s_hits2[0][threadIdx.y][threadIdx.x] = 0;
    const bool pending = A.pending_seq != 0u && A.ctl->nb_applied_seq != A.pending_seq;
    const uint32_t N = (uint32_t)A.N, REQ = (uint32_t)A.REQ;
    const uint32_t colorRange = CH == 1 ? 255u : 765u, descRange = CH == 1 ? 16u : 48u;
    uint32_t nz_acc = 0, sc_acc = 0, fg_acc = 0;
    __syncthreads();          // mbarriers initialised, tables staged, first hit plane cleared

    for(int it = 0; tile < ntiles; ++it, tile += (int)gridDim.x) {
        const int b = it & 1;
        uchar* s_buf = s_bufs[b];
        uchar* s_tile = s_buf + ST::O_IMG;
        const Col (*s_pcol)[ST::BW_COL] = (const Col (*)[ST::BW_COL])(s_buf + ST::O_PCOL);
        const Desc (*s_pdesc)[ST::BW_DESC] = (const Desc (*)[ST::BW_DESC])(s_buf + ST::O_PDESC);
        const ushort (*s_int)[ST::BW_INT] = (const ushort (*)[ST::BW_INT])(s_buf + ST::O_INT);
        const uchar (*s_own)[TILE_W] = (const uchar (*)[TILE_W])(s_buf + ST::O_OWN);
        const float (*s_R)[TILE_W] = (const float (*)[TILE_W])(s_buf + ST::O_R);
        const Rec (*s_smp)[TILE_H][TILE_W] = (const Rec (*)[TILE_H][TILE_W])(s_buf + ST::O_SMP);
        uint32_t (*s_hits)[TILE_W] = s_hits2[b];
        const int x0 = tx * TILE_W, y0 = ty * TILE_H;
        // next tile: its boxes go to the other buffer, which every thread finished reading before the barrier that closed the previous tile
        int ntx = tx + step_x, nty = ty + step_y;
        if(ntx >= tiles_x) { ntx -= tiles_x; ++nty; }
        if(tid == 0 && tile + (int)gridDim.x < ntiles) issue(ntx, nty, b ^ 1);
        if(!A.use_tma) { // frame pitch not TMA-compatible: cooperative byte copy of the input tile only
            const int rowbytes = (TILE_W + 2 * HALO) * CH;
            for(int i = tid; i < rowbytes * TILE_ROWS; i += TILE_W * TILE_H) {
                const int r = i / rowbytes, bb = i - r * rowbytes;
                const int gy = y0 - HALO + r, gb = (x0 - HALO) * CH + bb;
                uchar v = 0;
                if(gy >= 0 && gy < A.H && gb >= 0 && gb < A.W * CH) v = A.img[(size_t)gy * A.ipitch + gb];
                s_tile[r * PITCH + tile_shift(CH) + bb] = v;
            }
            __syncthreads();
        }

        const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
        const bool in_img = (x < A.W) && (y < A.H);
        const int wi = y * A.WW + (x >> 5);
        const uint32_t lane_bit = 1u << (x & 31);
        uint32_t w_roi = 0, w_unst = 0;
        if(y < A.H && (x >> 5) < A.WW) { w_roi = A.roi_bits[wi]; w_unst = A.unstable_bits[wi]; }
        const bool active = in_img && (w_roi & lane_bit);
        const size_t pix = (size_t)y * A.Wp + x;
        uint2 cbx = cbox_full();
        if(A.cbox && in_img) cbx = A.cbox[pix];   // (in flight while the tile's boxes land)

        mbar_wait(&s_bar[b], (uint32_t)((it >> 1) & 1));     // every TMA box of this tile has landed (async proxy -> visible after the phase flips)

        if(pending) {
            // scatter inside the CTA: every intent of the tile + halo marks its target pixel (smem atomics) instead of every target
            // scanning its 25 possible sources. Out-of-image entries of the box are zero-filled by TMA: they are skipped by position.
            static_assert(IW % 2 == 0, "intent pairs");
            for(int t = tid; t < (IW / 2) * IH; t += TILE_W * TILE_H) {
                const int r = t / (IW / 2), cp = t - r * (IW / 2);
                const int gx = x0 - HALO + 2 * cp, gy = y0 - HALO + r;
                if(gy < 0 || gy >= A.H) continue;
                const uint32_t pair = *(const uint32_t*)&s_int[r][2 * cp + ST::SH_INT];
#pragma unroll
                for(int e = 0; e < 2; ++e) {
                    const int cc = 2 * cp + e;
                    const uint32_t itw = e ? (pair >> 16) : (pair & 0xFFFFu);
                    if(gx + e < 0 || gx + e >= A.W || itw == NO_INTENT) continue; // padding columns of the plane are never written
                    const int code = (int)(itw >> 8), oy = code / 5, ox = code - oy * 5; // target = source + (ox-2, oy-2)
                    const int ttx = cc - 2 * HALO + ox, tty = r - 2 * HALO + oy;         // target inside the 32 x TILE_H core?
                    if(ttx >= 0 && ttx < TILE_W && tty >= 0 && tty < TILE_H) atomicOr(&s_hits[tty][ttx], 1u << (24 - code));
                }
            }
            __syncthreads();
        }

        float R = 0.f;
        Col lc = Col(), pre_c0 = Col(), pre_c1 = Col();
        Desc ld = Desc(), pre_d0 = Desc(), pre_d1 = Desc();
        if(in_img) {
            R = s_R[threadIdx.y][threadIdx.x];
            lc = s_pcol[threadIdx.y + HALO][threadIdx.x + HALO + ST::SH_COL];
            ld = s_pdesc[threadIdx.y + HALO][threadIdx.x + HALO + ST::SH_DESC];
            { const Rec r0 = s_smp[0][threadIdx.y][threadIdx.x]; pre_c0 = rec_col(r0); pre_d0 = rec_desc(r0); }
            if(A.N > 1) { const Rec r1 = s_smp[1][threadIdx.y][threadIdx.x]; pre_c1 = rec_col(r1); pre_d1 = rec_desc(r1); }
        }
        // apply the pending sample writes aimed at this pixel before anything reads the model: own stores are visible to own loads and
        // the two staged samples are patched in registers
        if(pending && in_img) {
            // the pixel's own stochastic update of the previous frame first (the reference writes it inside its pixel loop, the queued
            // neighbour writes come after the loop and override it): the record is the previous frame's colour / descriptors of this
            // very pixel, which the scan holds anyway. Doing the scattered store here, where it overlaps ~1000 instructions of
            // arithmetic, costs nothing measurable; at the end of the feedback kernel it cost 27 us per 1080p frame.
            const uint2 cbx_in = cbx;
            const uint32_t own = s_own[threadIdx.y][threadIdx.x];
            if(own != 0xFFu) {
#ifndef LVB_EXP_NO_OWN_WRITE
                ((Rec*)A.bg)[(size_t)own * A.plane + pix] = rec_make(lc, ld);
#endif
                cbox_grow(cbx, col_as_u32_(lc));
                if(own == 0u) { pre_c0 = lc; pre_d0 = ld; }
                if(own == 1u) { pre_c1 = lc; pre_d1 = ld; }
            }
            // neighbour writes in raster order of their source (ascending bit index): the source's record comes from the staged tiles
            uint32_t hits = s_hits[threadIdx.y][threadIdx.x];
            while(hits) {
                const int i = __ffs(hits) - 1;
                hits &= hits - 1;
                const int r = i / 5, k = i - r * 5;
                const uint32_t slot = s_int[threadIdx.y + r][threadIdx.x + k + ST::SH_INT] & 0xFFu;
                const Col c_ = s_pcol[threadIdx.y + r][threadIdx.x + k + ST::SH_COL]; const Desc d_ = s_pdesc[threadIdx.y + r][threadIdx.x + k + ST::SH_DESC];
#ifndef LVB_EXP_NO_NB_WRITE
                ((Rec*)A.bg)[(size_t)slot * A.plane + pix] = rec_make(c_, d_);
#endif
                if(slot == 0u) { pre_c0 = c_; pre_d0 = d_; }
                if(slot == 1u) { pre_c1 = c_; pre_d1 = d_; }
                cbox_grow(cbx, col_as_u32_(c_));
            }
            if(A.cbox && (cbx.x != cbx_in.x || cbx.y != cbx_in.y)) A.cbox[pix] = cbx;
        }
        s_hits2[b ^ 1][threadIdx.y][threadIdx.x] = 0; // the other hit plane, for the next tile (last read before the barrier that closed the previous tile)

        bool is_fg = false, nonzero = false;
        uint32_t cur[CH], intra[CH];
        uint32_t good = 0, s = 0, minDesc = descRange, minSum = colorRange;
        Col cur_pack = Col(); Desc intra_pack = Desc();
        Lookup16 Lk[CH];
        uint32_t thrC_ = 0, thrD_ = 0;

        if(active) {
            const bool unstable_old = (w_unst & lane_bit) != 0;
            // thresholds (SuBSENSE.cpp:222-223 / :355-359)
            uint32_t thrC = (uint32_t)(__fsub_rn(__fmul_rn(R, (float)A.min_color), (float)(unstable_old ? 0 : A.min_color / 5)));
            if(CH == 1) thrC >>= 1;
            const uint32_t thrD = (1u << (uint32_t)floorf(__fadd_rn(R, 0.5f))) + (uint32_t)A.desc_off + (unstable_old ? (uint32_t)A.desc_off : 0u);

            const int sy = threadIdx.y + HALO;
            Lookup16 (&L)[CH] = Lk;
            {
                const Window5<CH> Wn = lbsp_window_smem<CH>(s_tile, PITCH, sy, tile_shift(CH) + (int)threadIdx.x * CH);
#pragma unroll
                for(int c = 0; c < CH; ++c) {
                    L[c] = lbsp_lookup_window<CH>(Wn, c);
                    cur[c] = win_center<CH>(Wn, c);
                    intra[c] = lbsp_threshold<T7>(L[c], cur[c], s_lut[cur[c]]);
                }
            }
            if constexpr (CH == 1) { cur_pack = (uchar)cur[0]; intra_pack = (ushort)intra[0]; }
            else { cur_pack = cur[0] | (cur[1] << 8) | (cur[2] << 16); intra_pack = make_uint2(intra[0] | (intra[1] << 16), intra[2]); }

            // no sample can pass the colour gates (ColorBox above): foreground with zero matches, as if all N samples had been scanned
            if(A.cbox && cbox_excludes<CH>(cbx, col_as_u32_(cur_pack), CH == 1 ? thrC : (thrC * 3u) >> 1, thrC * 3u)) s = N;
            // samples 0 and 1 were staged with the tile
            uint32_t d_, s_;
            if(good < REQ && s < N) { if(subsense_test_sample<CH, T7>(L, cur, intra, pre_c0, pre_d0, thrC, thrD, s_lut, d_, s_)) { minDesc = min(minDesc, d_); minSum = min(minSum, s_); ++good; } ++s; }
            if(good < REQ && s < N) { if(subsense_test_sample<CH, T7>(L, cur, intra, pre_c1, pre_d1, thrC, thrD, s_lut, d_, s_)) { minDesc = min(minDesc, d_); minSum = min(minSum, s_); ++good; } ++s; }
            thrC_ = thrC; thrD_ = thrD;
        }
        // still undecided after the two staged samples: the pixel goes to the tail passes. One atomic per warp reserves the work-list
        // entries; it is issued HERE and its result used after the epilogue's stores, so the round trip overlaps them.
        const bool undecided = active && good < REQ && s < N;
        const uint32_t um = __ballot_sync(0xFFFFFFFFu, undecided);
        uint32_t base = 0;
        if(um && threadIdx.x == 0) base = atomicAdd(&A.ctl->wl_count, (uint32_t)__popc(um));

        if(active) {
            is_fg = good < REQ && s >= N; // an undecided pixel (s < N) is classified by the tail passes, which set its raw bit
            // distance to the previous frame (:254-255 / :396-397); the 3-channel L1 wraps in uint8 (quirk Q1)
            uint32_t lastL1 = 0, lastHd = 0;
#pragma unroll
            for(int c = 0; c < CH; ++c) {
                const uint32_t bb = col_get(lc, c);
                lastL1 += cur[c] > bb ? cur[c] - bb : bb - cur[c];
                lastHd += __popc(desc_get(ld, c) ^ intra[c]);
            }
            if(CH != 1) lastL1 &= 0xFFu;
            uint32_t pc = 0;
#pragma unroll
            for(int c = 0; c < CH; ++c) pc += __popc(intra[c]);
            nonzero = pc >= (CH == 1 ? 2u : 4u);
            A.hand[pix] = make_uint2(minSum | (minDesc << 16), good | (lastL1 << 16) | (lastHd << 24));
            ((Col*)A.last_color)[pix] = cur_pack;
            ((Desc*)A.last_desc)[pix] = intra_pack;
        }
        const uint32_t b_raw = __ballot_sync(0xFFFFFFFFu, is_fg);
        const uint32_t b_nz = __ballot_sync(0xFFFFFFFFu, nonzero);
        if(y < A.H && (x >> 5) < A.WW && threadIdx.x == 0) A.raw_bits[wi] = b_raw;
        nz_acc += __popc(b_nz); fg_acc += __popc(b_raw); sc_acc += s;   // (nz / fg: identical on every lane; lane 0 publishes them)

        if(um) { // work-list append
            base = __shfl_sync(0xFFFFFFFFu, base, 0);
            if(undecided) {
                const uint32_t e = base + (uint32_t)__popc(um & ((1u << threadIdx.x) - 1u));
                uint32_t f[X::FIELDS];
#pragma unroll
                for(int i = 0; i < X::FIELDS; ++i) f[i] = 0u;
#pragma unroll
                for(int c = 0; c < CH; ++c)
#pragma unroll
                    for(int q = 0; q < 4; ++q) f[X::LOOK + 4 * c + q] = Lk[c].w[q];
                if constexpr (CH == 1) { f[X::CUR] = cur[0]; f[X::INTRA] = intra[0]; }
                else { f[X::CUR] = cur_pack; f[X::INTRA] = intra_pack.x; f[X::INTRA + 1] = intra_pack.y; }
                f[X::THRC] = thrC_; f[X::THRD] = thrD_; f[X::PIX] = (uint32_t)pix;
                f[X::STATE] = wl_state_pack(minSum, minDesc, good, s);
                uint4* w = (uint4*)(A.wl_ctx + (size_t)e * X::FIELDS);
#pragma unroll
                for(int v = 0; v < X::VEC; ++v) w[v] = make_uint4(f[4 * v], f[4 * v + 1], f[4 * v + 2], f[4 * v + 3]);
            }
        }
        tx = ntx; ty = nty;
        __syncthreads(); // every thread is done with this tile's buffer and hit plane: they are refilled two / one iterations from now
    }

    if(A.collect_stats) {
#pragma unroll
        for(int o = 16; o > 0; o >>= 1) sc_acc += __shfl_xor_sync(0xFFFFFFFFu, sc_acc, o);
    }
    // the last warp of the CTA to get here publishes the CTA's counters (the loop's closing barrier keeps the warps within one tile of each other)
    if(threadIdx.x == 0) {
        atomicAdd(&s_cnt[0], nz_acc);
        if(A.collect_stats) { atomicAdd(&s_cnt[1], sc_acc); atomicAdd(&s_cnt[3], fg_acc); }
        __threadfence_block();
        if(atomicAdd(&s_cnt[4], 1u) == (uint32_t)TILE_H - 1u) {
            __threadfence_block();
            const uint32_t nz = atomicAdd(&s_cnt[0], 0u);
            if(nz) atomicAdd(&A.ctl->nonzero_count, nz);
            if(A.collect_stats) {
                atomicAdd(&A.ctl->stat_scanned, (unsigned long long)atomicAdd(&s_cnt[1], 0u));
                atomicAdd(&A.ctl->stat_fg, (unsigned long long)atomicAdd(&s_cnt[3], 0u));
            }
        }
    }
}

/// Tail passes of the sample-consensus scan (SuBSENSE.cpp:229-253 / :367-395 past the second sample): one work-list entry per
/// LANE, scanned sequentially exactly like the reference's loop ("while good < REQ and s < N"), with the sample records of a batch
/// of B consecutive samples in flight together. Pass 1 takes every entry up to sample `s_limit` (8: more than 98 % of the
/// background pixels are decided by then); the survivors (foreground, which scans all N, and a few deep background pixels) go on
/// list 2 by index and are finished by pass 2, where almost every lane runs the full N samples, so warps stay converged.
/// A decided entry completes the scan kernel's outputs: min distances + match count in the hand-off word, raw bit if foreground.
#ifndef LVB_TAIL_PASS1_LIMIT
#define LVB_TAIL_PASS1_LIMIT 8
#endif
constexpr uint32_t TAIL_PASS1_LIMIT = LVB_TAIL_PASS1_LIMIT; // pass 1 scans samples 2 .. LIMIT-1
#ifndef LVB_TAIL1_MINB
#define LVB_TAIL1_MINB 7
#endif
#ifndef LVB_TAIL1_B
#define LVB_TAIL1_B 3
#endif
constexpr int TAIL1_B = LVB_TAIL1_B;   // pass 1: sample records per batch (in registers)
#ifndef LVB_TAIL2_MINB
#define LVB_TAIL2_MINB 4
#endif
#ifndef LVB_TAIL2_B
#define LVB_TAIL2_B 8
#endif
constexpr int TAIL1_MINB = LVB_TAIL1_MINB, TAIL2_MINB = LVB_TAIL2_MINB, TAIL2_B = LVB_TAIL2_B; // resident CTAs per SM (launch bounds) of the two passes; pass-2 stage depth
struct TailPassArgs {
    int Wp, WW, N, REQ;
    size_t plane;
    const void* bg;
    uint32_t* wl_ctx; uint32_t wl_cap;
    const uint32_t* in_idx;     // nullptr: entries 0..count-1 of the work-list; else the entries named by this index list
    const uint32_t* in_count;
    uint32_t* cursor;           // work distribution: every warp pulls chunks of 32 entries (FrameCtl::wl_cursor / wl2_cursor)
    uint32_t* out_idx; uint32_t* out_count;   // survivors (nullptr: none expected, s_limit >= N)
    uint32_t s_limit;
    uint2* hand; uint32_t* raw_bits;
    const uchar* lut;
    FrameCtl* ctl; int collect_stats;
};
/// STREAM = false (pass 1): the B records of a batch are loaded into registers. STREAM = true (pass 2, where nearly every entry scans
/// all N samples): the records stream through a per-lane double buffer in shared memory filled by cp.async, batch j+1 in flight while
/// batch j is tested, so the ~40 dependent DRAM round trips of a foreground pixel collapse into one pipeline.
template<int CH, bool T7, int B, int MINB, bool STREAM>
__global__ void __launch_bounds__(128, MINB) subsense_tail_pass(const TailPassArgs A) {
    typedef typename Pack<CH>::Col Col;
    typedef typename Pack<CH>::Desc Desc;
    typedef typename Pack<CH>::Rec Rec;
    typedef WlCtx<CH> X;
    __shared__ uchar s_lut[256];
    __shared__ __align__(16) Rec s_rec[STREAM ? 2 : 1][STREAM ? B : 1][STREAM ? 128 : 1];
    pdl_launch_dependents();
    for(int i = threadIdx.x; i < 256; i += blockDim.x) s_lut[i] = A.lut[i];   // written by the previous frame's tail only
    __syncthreads();
    pdl_wait();
    const uint32_t count = *A.in_count;
    const uint32_t N = (uint32_t)A.N, REQ = (uint32_t)A.REQ, s_end = min(N, A.s_limit);
    const uint32_t lane = threadIdx.x & 31u, tid = threadIdx.x;
    uint32_t scanned_acc = 0, fg_acc = 0;
    for(;;) {
        uint32_t e0 = 0;
        if(lane == 0) e0 = atomicAdd(A.cursor, 32u);
        e0 = __shfl_sync(0xFFFFFFFFu, e0, 0);
        if(e0 >= count) break;
        const uint32_t e = e0 + lane;
        const bool have = e < count;
        bool survivor = false;
        uint32_t idx = 0;
        if(have) {
            idx = A.in_idx ? A.in_idx[e] : e;
            uint32_t* w = A.wl_ctx + (size_t)idx * X::FIELDS;
            uint32_t f[X::FIELDS];
#pragma unroll
            for(int v = 0; v < X::VEC; ++v) { const uint4 t = ((const uint4*)w)[v]; f[4 * v] = t.x; f[4 * v + 1] = t.y; f[4 * v + 2] = t.z; f[4 * v + 3] = t.w; }
            Lookup16 L[CH];
            uint32_t cur[CH], intra[CH];
#pragma unroll
            for(int c = 0; c < CH; ++c)
#pragma unroll
                for(int q = 0; q < 4; ++q) L[c].w[q] = f[X::LOOK + 4 * c + q];
            const uint32_t cp = f[X::CUR];
            if constexpr (CH == 1) { cur[0] = cp; intra[0] = f[X::INTRA]; }
            else {
#pragma unroll
                for(int c = 0; c < CH; ++c) cur[c] = (cp >> (8 * c)) & 0xFFu;
                intra[0] = f[X::INTRA] & 0xFFFFu; intra[1] = f[X::INTRA] >> 16; intra[2] = f[X::INTRA + 1];
            }
            const uint32_t thrC = f[X::THRC], thrD = f[X::THRD], pix = f[X::PIX];
            const uint32_t st = f[X::STATE];
            uint32_t minSum = st & 0x3FFu, minDesc = (st >> 10) & 0x3Fu, good = (st >> 16) & 0xFFu, s = st >> 24;
            const uint32_t s0 = s;
            const Rec* bgp = (const Rec*)A.bg + pix;
            if constexpr (STREAM) {
                auto issue = [&](int stage, uint32_t from) {
#pragma unroll
                    for(int j = 0; j < B; ++j)
                        if(from + j < s_end) {
                            if constexpr (sizeof(Rec) == 16) cp_async16(&s_rec[stage][j][tid], bgp + (size_t)(from + j) * A.plane);
                            else cp_async4_zfill(&s_rec[stage][j][tid], bgp + (size_t)(from + j) * A.plane, true);
                        }
                    cp_async_commit();
                };
                int stage = 0;
                issue(0, s);
                while(good < REQ && s < s_end) {
                    issue(stage ^ 1, s + B);   // next batch in flight while this one is tested
                    cp_async_wait<1>();        // this batch has landed (every lane reads only what it copied itself)
#pragma unroll
                    for(int j = 0; j < B; ++j) {
                        if(good < REQ && s < s_end) {
                            const Rec r = s_rec[stage][j][tid];
                            uint32_t d_, s_;
                            if(subsense_test_sample<CH, T7>(L, cur, intra, rec_col(r), rec_desc(r), thrC, thrD, s_lut, d_, s_)) { minDesc = min(minDesc, d_); minSum = min(minSum, s_); ++good; }
                            ++s;
                        }
                    }
                    stage ^= 1;
                }
                cp_async_wait<0>();            // nothing of this entry may land in the buffers once the next entry uses them
            } else {
                while(good < REQ && s < s_end) {
                    Rec r[B];
#pragma unroll
                    for(int j = 0; j < B; ++j) if(s + j < s_end) r[j] = bgp[(size_t)(s + j) * A.plane];
#pragma unroll
                    for(int j = 0; j < B; ++j) {
                        if(good < REQ && s < s_end) {
                            uint32_t d_, s_;
                            if(subsense_test_sample<CH, T7>(L, cur, intra, rec_col(r[j]), rec_desc(r[j]), thrC, thrD, s_lut, d_, s_)) { minDesc = min(minDesc, d_); minSum = min(minSum, s_); ++good; }
                            ++s;
                        }
                    }
                }
            }
            scanned_acc += s - s0;
            if(good < REQ && s < N) { // not decided within this pass
                survivor = true;
                w[X::STATE] = wl_state_pack(minSum, minDesc, good, s);
            } else {
                uint2* h = A.hand + pix;
                h->x = minSum | (minDesc << 16);
                *(ushort*)&h->y = (ushort)good; // low half of (good | lastL1 << 16 | lastHd << 24)
                if(good < REQ) {
                    const uint32_t px = pix % (uint32_t)A.Wp, py = pix / (uint32_t)A.Wp;
                    atomicOr(A.raw_bits + (size_t)py * A.WW + (px >> 5), 1u << (px & 31u));
                    ++fg_acc;
                }
            }
        }
        const uint32_t sm = __ballot_sync(0xFFFFFFFFu, survivor);
        if(sm && A.out_idx) {
            uint32_t base = 0;
            if(lane == 0) base = atomicAdd(A.out_count, (uint32_t)__popc(sm));
            base = __shfl_sync(0xFFFFFFFFu, base, 0);
            if(survivor) A.out_idx[base + (uint32_t)__popc(sm & ((1u << lane) - 1u))] = idx;
        }
    }
    if(A.collect_stats) {
#pragma unroll
        for(int o = 16; o > 0; o >>= 1) { scanned_acc += __shfl_xor_sync(0xFFFFFFFFu, scanned_acc, o); fg_acc += __shfl_xor_sync(0xFFFFFFFFu, fg_acc, o); }
        if(lane == 0) {
            if(scanned_acc) atomicAdd(&A.ctl->stat_scanned, (unsigned long long)scanned_acc);
            if(fg_acc) atomicAdd(&A.ctl->stat_fg, (unsigned long long)fg_acc);
        }
    }
}

/// frame tail (SuBSENSE.cpp:555-611): LUT +-1 adaptation, frame-level reset / learning-rate caps, next-frame factors.
/// Executed by one warp: the last one of the last CTA of the feedback kernel.
#ifndef LVB_SPIN_LIMIT
#define LVB_SPIN_LIMIT 4000000u   // x 500 ns: the frame tail waits at most ~2 s for the frame's final mask before it drops a model reset
#endif
struct TailArgs {
    FrameCtl* ctl; uchar* lut;
    float rel; int lbsp_off; int min_color; int avg_samples; int N; int dsW, dsH;
    uint64_t seed;
    uint32_t wait_seq;         // a requested reset holds this warp until FrameCtl::chain_done reaches this frame (see below)
};
__device__ __noinline__ void subsense_tail_warp(const TailArgs& A, int lane) {
    FrameCtl* ctl = A.ctl;
    int dir = 0;
    if(lane == 0) {
        const float ratio = __fdiv_rn((float)ctl->nonzero_count, (float)ctl->roi_count);
        const float last = ctl->last_nonzero_ratio;
        dir = (ratio < 0.1f && last < 0.1f) ? -1 : (ratio > 0.5f && last > 0.5f) ? 1 : 0;
        ctl->last_nonzero_ratio = ratio;
        ctl->nonzero_count = 0;
        ctl->wl_count = 0; ctl->wl2_count = 0; ctl->wl_cursor = 0; ctl->wl2_cursor = 0; // this frame's scan work-lists are consumed
        ctl->do_refresh = 0; ctl->set_T_one = 0;
        if(ctl->lr_scaling) {
            const float diff_ratio = __fdiv_rn((float)ctl->tot_color_diff, (float)(A.dsW * A.dsH));
            const uint32_t thr = (uint32_t)A.min_color / 2u;
            if(ctl->auto_reset) {
                if(ctl->frames_since_reset > 1000u) ctl->auto_reset = 0;
                else if(diff_ratio >= (float)thr && ctl->cooldown == 0) {
                    ctl->frames_since_reset = 0;
                    // refreshModel(0.1f): (size_t)(0.1f*N) slots starting at a random position
                    ctl->do_refresh = 1; ctl->set_T_one = 1; ctl->refresh_force = 0;
                    ctl->refresh_count = (uint32_t)__fmul_rn(0.1f, (float)A.N);
                    const uint32_t epoch = ctl->refresh_epoch;
                    ctl->refresh_start = philox_draw(A.seed, epoch, 0, 0, DOM_REFRESH_START) % (uint32_t)A.N;
                    ctl->cooldown = (uint32_t)A.avg_samples / 4u;
                } else ctl->frames_since_reset += 1;
            } else if(diff_ratio >= (float)(thr * 2u)) {
                ctl->frames_since_reset = 0;
                ctl->auto_reset = 1;
            }
            if(diff_ratio >= (float)(thr / 2u)) {
                const int sh = (int)__fdiv_rn(diff_ratio, 2.0f);
                ctl->t_lower = (float)max(sh < 31 ? (2 >> sh) : 0, 1);
                ctl->t_upper = (float)max(sh < 31 ? (256 >> sh) : 0, 1);
            } else { ctl->t_lower = 2.0f; ctl->t_upper = 256.0f; }
            if(ctl->cooldown > 0) ctl->cooldown -= 1;
            ctl->tot_color_diff = 0;
        }
        // next frame
        const uint32_t f = ctl->frame_idx + 1;
        ctl->frame_idx = f;
        ctl->aLT = __fdiv_rn(1.0f, (float)min(f, (uint32_t)A.avg_samples));
        ctl->aST = __fdiv_rn(1.0f, (float)min(f, (uint32_t)A.avg_samples / 4u));
        ctl->blocks_done = 0;
        // A requested reset (rare: scene change) resamples the model from pixels that are background in THIS frame's final mask,
        // which the mask stream is still producing. This one warp (one resident CTA per instance, so any number of instances may
        // wait at once without starving the mask kernels they wait for) holds the feedback kernel open until the mask is complete;
        // the refresh kernel behind it in the stream then needs no synchronisation of its own.
        // The wait is bounded (~2 s): it relies on the mask stream making progress beside this kernel, which holds whenever the mask
        // chain was enqueued first (it always is, so even CUDA_LAUNCH_BLOCKING=1 finds the flag set), but is not a CUDA guarantee
        // (time-sliced contexts). On a timeout the request is dropped and FrameCtl::spin_timeout names the frame: the host reports it
        // as an error at its next synchronisation point instead of the device hanging.
        if(ctl->do_refresh && A.wait_seq) {
            volatile uint32_t* flag = &ctl->chain_done;
            uint32_t spins = 0;
            while((int32_t)(*flag - A.wait_seq) < 0 && spins < LVB_SPIN_LIMIT) { __nanosleep(500); ++spins; }
            if((int32_t)(*flag - A.wait_seq) < 0) { ctl->do_refresh = 0; ctl->set_T_one = 0; ctl->spin_timeout = A.wait_seq; }
            __threadfence();
        }
    }
    dir = __shfl_sync(0xFFFFFFFFu, dir, 0);
    for(int t = lane; t < 256; t += 32) {
        if(dir < 0) {
            const float lo = fminf(fmaxf(rintf(__fadd_rn((float)A.lbsp_off, ceilf(__fdiv_rn(__fmul_rn((float)t, A.rel), 4.0f)))), 0.f), 255.f);
            if((float)A.lut[t] > lo) A.lut[t] -= 1;
        } else if(dir > 0) {
            const float hi = fminf(fmaxf(rintf(__fadd_rn((float)A.lbsp_off, __fmul_rn(255.0f, A.rel))), 0.f), 255.f);
            if((float)A.lut[t] < hi) A.lut[t] += 1;
        }
    }
}

#ifndef FEEDBACK_MIN_BLOCKS
#define FEEDBACK_MIN_BLOCKS 5
#endif
constexpr int FB_H = 8;                       // tile height (warps per CTA) of the feedback kernel
constexpr int FB_CTAS_PER_SM = FEEDBACK_MIN_BLOCKS;
#ifndef LVB_FB_GRID_CTAS
#define LVB_FB_GRID_CTAS FEEDBACK_MIN_BLOCKS
#endif
constexpr int FB_GRID_CTAS_PER_SM = LVB_FB_GRID_CTAS;   // grid of the multi-tile feedback kernel, in CTAs per SM

/// one 32x8 tile of per-pixel state staged in shared memory by cp.async, one block per warp row (so a warp only ever reads what
/// its own lanes copied: the tile loop needs __syncwarp, never __syncthreads)
template<int CH> struct FbStage {
    float4 maps[FB_H][64];                    // (T,R,v,Dlast | DminLT,DminST,rawLT,rawST) x 32 px
    float2 fin[FB_H][32];
    uint2 hand[FB_H][32];
    uint32_t words[FB_H][20];                 // roi | blinks | lastfg | - | previous frame's ghost bits: rows y-2..y+2 x words wi-1..wi+1 | -
};

/// The kernel does not touch the sample model: the pixel's own stochastic update is QUEUED (own_slot plane, like the neighbour
/// write in the intents plane) and stored by the next frame's scan kernel, which holds the record anyway.
/// Persistent: a CTA walks 32x8 tiles with stride gridDim.x. The state of tile i+1 is copied global -> shared (cp.async, no
/// registers held) while tile i is computed, so the ~1 us DRAM round trip that used to open every CTA's life is hidden behind
/// the previous tile's arithmetic, and the small division / modulo tables are staged once per CTA instead of once per tile.
template<int CH>
__global__ void __launch_bounds__(32 * FB_H, FEEDBACK_MIN_BLOCKS)
subsense_feedback(const SubArgs A, const TailArgs TA) {
    typedef typename Pack<CH>::Col Col;
    typedef typename Pack<CH>::Desc Desc;
    typedef typename Pack<CH>::Rec Rec;
    __shared__ uint32_t s_cnt[2];                 // writes | warps done
    __shared__ CtlSlice s_ctl;
    // small read-only tables behind data-dependent indices (T(x), the hand-off distances): staged in shared memory so that the
    // look-ups cost a fixed ~25 cycles instead of an L1 miss in the middle of the dependency chain (L1 is streamed through by
    // the feedback maps). [0,257): floor(2^32/n) ; then i/colorRange ; then i/descRange
    constexpr int NCOL = (CH == 1 ? 255 : 765) + 1, NDES = (CH == 1 ? 16 : 48) + 1;
    __shared__ uint32_t s_magic[257];
    __shared__ float s_divc[NCOL], s_divd[NDES];
    __shared__ __align__(16) FbStage<CH> s_stage[2];
    static_assert(sizeof(FbStage<CH>) % 16 == 0, "stage alignment");

    const int lane = threadIdx.x, warp = threadIdx.y;
    const int tid = warp * TILE_W + lane;
    const int tiles_x = A.Wp / TILE_W, ntiles = tiles_x * ((A.H + FB_H - 1) / FB_H);

    // tile coordinates advance incrementally (no integer division per tile): (tx, ty) += (step_x, step_y) with a carry
    const int step_y = (int)gridDim.x / tiles_x, step_x = (int)gridDim.x - step_y * tiles_x;
    auto advance = [&](int& tx, int& ty) { tx += step_x; ty += step_y; if(tx >= tiles_x) { tx -= tiles_x; ++ty; } };
    // which flag word this lane stages for its warp: roi | blinks | lastfg | previous frame's ghost bits (rows y-2..y+2, words wi-1..wi+1)
    const uint32_t* w_src = lane == 0 ? A.roi_bits : lane == 1 ? A.blinks_bits : lane == 2 ? A.lastfg_bits : A.ghost_prev;
    const int w_dy = lane >= 3 ? (lane - 3) / 3 - HALO : 0, w_dx = lane >= 3 ? (lane - 3) % 3 - 1 : 0, w_slot = lane < 3 ? lane : lane + 1;
    auto prefetch = [&](int tx, int ty, FbStage<CH>& S) {
        const int x0 = tx * TILE_W, y = ty * FB_H + warp;
        if(y < A.H) { // whole 32-px row segments: the planes are Wp (a multiple of 32) wide, padding columns are never used
            const size_t rowpix = (size_t)y * A.Wp + x0;
            const char* g_maps = (const char*)(A.maps + rowpix * 2);
            cp_async16((char*)&S.maps[warp][0] + 16 * lane, g_maps + 16 * lane);
            cp_async16((char*)&S.maps[warp][0] + 16 * (lane + 32), g_maps + 16 * (lane + 32));
            if(lane < 16) cp_async16((char*)&S.fin[warp][0] + 16 * lane, (const char*)(A.fin + rowpix) + 16 * lane);
            else cp_async16((char*)&S.hand[warp][0] + 16 * (lane - 16), (const char*)(A.hand + rowpix) + 16 * (lane - 16));
            if(lane < 18) {
                const int gy = y + w_dy, gw = tx + w_dx;
                const bool ok = gy >= 0 && gy < A.H && gw >= 0 && gw < A.WW;
                cp_async4_zfill(&S.words[warp][w_slot], ok ? w_src + (size_t)gy * A.WW + gw : w_src, ok);
            }
        }
    };

    int tile = blockIdx.x;
    int tx = tile % tiles_x, ty = tile / tiles_x;
    if(tile < ntiles) prefetch(tx, ty, s_stage[0]);
    cp_async_commit();
    for(int i = tid; i < 257; i += TILE_W * FB_H) s_magic[i] = A.magic[i];
    for(int i = tid; i < NCOL; i += TILE_W * FB_H) s_divc[i] = A.div_color[i];
    if(tid < NDES) s_divd[tid] = A.div_desc[tid];
    if(tid < 2) s_cnt[tid] = 0;
    if(tid == 32 * FB_H - 1) {
        const FrameCtl* ctl = A.ctl;
        CtlSlice cs;
        cs.aLT = ctl->aLT; cs.aST = ctl->aST; cs.t_lower = ctl->t_lower; cs.t_upper = ctl->t_upper;
        cs.frame = ctl->frame_idx; cs.cooldown = ctl->cooldown; cs.use3x3 = ctl->use3x3; cs.pad = 0;
        s_ctl = cs;
    }
    __syncthreads(); // the only CTA-wide barrier: tables + FrameCtl slice
    const uint32_t N = (uint32_t)A.N, REQ = (uint32_t)A.REQ;
    const uint32_t colorRange = CH == 1 ? 255u : 765u, descRange = CH == 1 ? 16u : 48u;
    const float aLT = s_ctl.aLT, aST = s_ctl.aST, t_lower = s_ctl.t_lower, t_upper = s_ctl.t_upper;
    const uint32_t frame = s_ctl.frame, cooldown = s_ctl.cooldown, use3x3 = s_ctl.use3x3;
    uint32_t writes_acc = 0;
    // factors of the folded final-mask EMAs, uniform over the frame: (1 - e) and 255 * ((1/255) * e), each rounded as in the per-pixel form
    const float eLT = __fdiv_rn(1.0f, (float)min(max(A.ema_frame, 1u), (uint32_t)A.avg_samples)), eST = __fdiv_rn(1.0f, (float)min(max(A.ema_frame, 1u), (uint32_t)A.avg_samples / 4u));
    const double ema_a1 = (double)__fsub_rn(1.0f, eLT), ema_a2 = (double)__fsub_rn(1.0f, eST);
    const double ema_b1 = __dmul_rn(255.0, __dmul_rn(1.0 / 255, (double)eLT)), ema_b2 = __dmul_rn(255.0, __dmul_rn(1.0 / 255, (double)eST));

    for(int it = 0; tile < ntiles; ++it, tile += gridDim.x) {
        const FbStage<CH>& S = s_stage[it & 1];
        const int next = tile + (int)gridDim.x;
        int ntx = tx, nty = ty;
        advance(ntx, nty);
        if(next < ntiles) prefetch(ntx, nty, s_stage[(it + 1) & 1]);
        cp_async_commit();
        cp_async_wait<1>();  // everything but the group just committed has landed: this tile's state is in shared memory
        __syncwarp();        // ... and visible to the other lanes of the warp (row segments are copied cooperatively)

        const int x0 = tx * TILE_W;
        const int x = x0 + lane, y = ty * FB_H + warp;
        const bool in_img = (x < A.W) && (y < A.H);
        const int wi = y * A.WW + (x >> 5);
        const uint32_t lane_bit = 1u << lane;
        uint32_t w_roi = 0, w_blink = 0, w_lastfg = 0;
        if(y < A.H) { w_roi = S.words[warp][0]; w_blink = S.words[warp][1]; w_lastfg = S.words[warp][2]; }
        const bool active = in_img && (w_roi & lane_bit);
        const size_t pix = (size_t)y * A.Wp + x;

        bool unstable_new = false, ghost_new = false, has_intent = false;
        uint32_t writes = 0;
        uint32_t intent = NO_INTENT; // queued neighbour write: (clamped relative target offset index) << 8 | slot
        uint32_t own = 0xFFu;        // queued own-sample write: slot (the next scan stores this frame's colour / descriptors there)

        if(active) {
            const float4 m0 = S.maps[warp][2 * lane], m1 = S.maps[warp][2 * lane + 1];
            float2 fin = S.fin[warp][lane];
            const uint2 hand = S.hand[warp][lane];
            const uint32_t minSum = hand.x & 0xFFFFu, minDesc = hand.x >> 16, good = hand.y & 0xFFFFu, lastL1 = (hand.y >> 16) & 0xFFu, lastHd = hand.y >> 24;
            const bool is_fg = good < REQ;
            float T = m0.x, R = m0.y, V = m0.z, Dlast = m0.w, DminLT = m1.x, DminST = m1.y, rawLT = m1.z, rawST = m1.w;
            const bool blink = (w_blink & lane_bit) != 0, lastfg = (w_lastfg & lane_bit) != 0;
            if(A.ema_frame) { // final-segmentation EMAs of the previous frame (:553-554): cv::addWeighted accumulates in double, rounds once
                fin.x = (float)__dadd_rn(__dmul_rn((double)fin.x, ema_a1), lastfg ? ema_b1 : 0.0);
                fin.y = (float)__dadd_rn(__dmul_rn((double)fin.y, ema_a2), lastfg ? ema_b2 : 0.0);
                A.fin[pix] = fin;
            }
            unstable_new = (R > 3.0f) || (__fsub_rn(rawLT, fin.x) > 0.1f) || (__fsub_rn(rawST, fin.y) > 0.1f);

            // D_last (:254-255 / :396-397)
            // i / colorRange and i / descRange come from 3 KB of host-tabulated IEEE quotients instead of four __fdiv_rn sequences per pixel
            const float normLast = __fmul_rn(__fadd_rn(s_divc[min(lastL1, (uint32_t)NCOL - 1u)], s_divd[min(lastHd, (uint32_t)NDES - 1u)]), 0.5f); // x/2 == x*0.5 exactly
            Dlast = __fadd_rn(__fmul_rn(Dlast, __fsub_rn(1.0f, aST)), __fmul_rn(normLast, aST));

            const uint32_t pixid = (uint32_t)(y * A.W + x);
#ifndef LVB_EXP_NO_PHILOX
            const uint4 rnd = philox_block(A.seed, frame, pixid, 0, DOM_APPLY);
#else
            const uint4 rnd = make_uint4((pixid * 2654435761u + frame) >> 1, (pixid * 40503u + frame * 7u) >> 1, (pixid ^ (frame * 97u)) >> 1, (pixid * 2246822519u + frame) >> 1);
#endif
            const float oneLT = __fsub_rn(1.0f, aLT), oneST = __fsub_rn(1.0f, aST);
            const float baseMin = __fmul_rn(__fadd_rn(s_divc[min(minSum, colorRange)], s_divd[min(minDesc, descRange)]), 0.5f);
            if(is_fg) { // foreground (:256-269 / :398-413)
                const float normMin = fminf(1.0f, __fadd_rn(baseMin, __fdiv_rn((float)(REQ - good), (float)REQ)));
                DminLT = __fadd_rn(__fmul_rn(DminLT, oneLT), __fmul_rn(normMin, aLT));
                DminST = __fadd_rn(__fmul_rn(DminST, oneST), __fmul_rn(normMin, aST));
                rawLT = __fadd_rn(__fmul_rn(rawLT, oneLT), aLT);
                rawST = __fadd_rn(__fmul_rn(rawST, oneST), aST);
                if(cooldown && (rnd.x % 2u) == 0) {
                    own = fast_mod(rnd.y, N, A.n_magic);
                    ++writes;
                }
            } else { // background (:270-301 / :414-450)
                DminLT = __fadd_rn(__fmul_rn(DminLT, oneLT), __fmul_rn(baseMin, aLT));
                DminST = __fadd_rn(__fmul_rn(DminST, oneST), __fmul_rn(baseMin, aST));
                rawLT = __fmul_rn(rawLT, oneLT);
                rawST = __fmul_rn(rawST, oneST);
                // x % LR, x % (LR/2+1): T(x) <= 256, so the magic numbers come from a 1 KB table (a fixed rate has them in the arguments)
                const uint32_t LR = A.lr_fixed ? A.lr_fixed : (uint32_t)ceilf(T);
                const uint32_t LR2 = LR / 2u + 1u;
                const bool tab = !A.lr_fixed && LR <= 256u;
                const uint32_t mg = A.lr_fixed ? A.lr_magic : s_magic[tab ? LR : 0u], mg2 = A.lr_fixed ? A.lr2_magic : s_magic[tab ? LR2 : 0u];
                const bool fastm = A.lr_fixed || tab;
                if((fastm ? fast_mod(rnd.x, LR, mg) : rnd.x % LR) == 0) {
                    own = fast_mod(rnd.y, N, A.n_magic);
                    ++writes;
                }
                const bool cur3 = use3x3 && !unstable_new;
                int dx, dy;
                neighbor_offset(cur3, rnd.z, dx, dy);
                const int nx = clampi(x + dx, 2, A.W - 3), ny = clampi(y + dy, 2, A.H - 3);
                const bool nb_ghost = (S.words[warp][4 + (ny - y + HALO) * 3 + (nx >> 5) - (x0 >> 5) + 1] >> (nx & 31)) & 1u;
                const uint32_t n_rand = rnd.w;
                const uint32_t nbLR = cur3 ? LR : LR2;
                if((fastm ? fast_mod(n_rand, nbLR, cur3 ? mg : mg2) : n_rand % nbLR) == 0 || (nb_ghost && (n_rand % (uint32_t)t_lower) == 0)) {
                    const uint32_t slot = fast_mod(fast_div(rnd.y, N, A.n_magic), N, A.n_magic); // (d1 / N) % N: one Philox block serves the whole pixel
                    // intent = (clamped relative target offset index) << 8 | slot ; offset index = (ty-y+2)*5 + (tx-x+2)
                    intent = (((ny - y + 2) * 5 + (nx - x + 2)) << 8) | slot;
                    has_intent = true;
                }
            }
            // T(x) (:302-311 / :451-460)
            const float dmin = fminf(DminLT, DminST), dmax = fmaxf(DminLT, DminST);
            if(lastfg || (dmin < 0.1f && is_fg)) {
                if(T < t_upper) T = __fadd_rn(T, __fdiv_rn(0.5f, __fmul_rn(dmax, V)));
            } else if(T > t_lower)
                T = __fsub_rn(T, __fdiv_rn(__fmul_rn(0.25f, V), dmax));
            if(T < t_lower) T = t_lower; else if(T > t_upper) T = t_upper;
            // v(x) (:312-318 / :461-467)
            if(dmax > 0.1f && blink) V = __fadd_rn(V, 1.0f);
            else if(V > 0.1f) {
                V = __fsub_rn(V, lastfg ? (0.1f / 4) : unstable_new ? (0.1f / 2) : 0.1f);
                if(V < 0.1f) V = 0.1f;
            }
            // R(x) (:319-325 / :468-474); std::pow(float,int) evaluates in double (Q7): the square is exact in fp64
            const double rr = (double)__fadd_rn(1.0f, __fmul_rn(dmin, 2.0f));
            if((double)R < __dmul_rn(rr, rr)) R = __fadd_rn(R, __fmul_rn(0.01f, __fsub_rn(V, 0.1f)));
            else {
                R = __fsub_rn(R, __fdiv_rn(0.01f, V));
                if(R < 1.0f) R = 1.0f;
            }
            ghost_new = (rawST > 0.995f) && (Dlast < 0.010f);

            A.maps[pix * 2] = make_float4(T, R, V, Dlast);
            A.maps[pix * 2 + 1] = make_float4(DminLT, DminST, rawLT, rawST);
            A.r_plane[pix] = R;
        }

        // warp-level packing of the per-pixel flags: one 32-bit mask word per warp row
        const uint32_t b_unst = __ballot_sync(0xFFFFFFFFu, unstable_new);
        const uint32_t b_ghost = __ballot_sync(0xFFFFFFFFu, ghost_new);
        if(in_img) { A.intents[pix] = (ushort)intent; A.own_slot[pix] = (uchar)own; } // every pixel, every frame: consumers scan the planes without a has-intent mask
        if(y < A.H && (x >> 5) < A.WW && lane == 0) { A.unstable_bits[wi] = b_unst; A.ghost_cur[wi] = b_ghost; }
        if(A.collect_stats) writes_acc += writes + (has_intent ? 1u : 0u);
        __syncwarp(); // every lane is done with this stage before the warp refills it (two iterations from now)
        tx = ntx; ty = nty;
    }
    cp_async_wait<0>();
    if(A.collect_stats) {
#pragma unroll
        for(int o = 16; o > 0; o >>= 1) writes_acc += __shfl_xor_sync(0xFFFFFFFFu, writes_acc, o);
        if(lane == 0) atomicAdd(&s_cnt[0], writes_acc);
    }
    // frame tail: the last warp of the CTA to get here takes the CTA's ticket, and the last CTA of the grid runs the tail with
    // that one warp (every other CTA has read its FrameCtl slice long before it took its ticket). No CTA-wide barrier: warps
    // retire as they finish instead of idling through the ticket's round trip.
    uint32_t last_warp = 0;
    if(lane == 0) {
        __threadfence_block();
        last_warp = atomicAdd(&s_cnt[1], 1u) == (uint32_t)FB_H - 1u;
    }
    last_warp = __shfl_sync(0xFFFFFFFFu, last_warp, 0);
    if(!last_warp) return;
    uint32_t last_cta = 0;
    if(lane == 0) {
        const uint32_t wr = atomicAdd(&s_cnt[0], 0u);
        if(A.collect_stats && wr) atomicAdd(&A.ctl->stat_writes, (unsigned long long)wr);
        // no device-wide fence: the tail only reads counters accumulated with atomics (by this grid) or by earlier kernels,
        // and what it rewrites in FrameCtl was read by every CTA before that CTA's first barrier
        last_cta = atomicAdd(&A.ctl->blocks_done, 1u) == gridDim.x - 1u;
    }
    last_cta = __shfl_sync(0xFFFFFFFFu, last_cta, 0);
    if(last_cta) {
        subsense_tail_warp(TA, lane);
    }
}

/// Phase B: apply the queued neighbour writes. One thread per TARGET pixel scans the intent words of the 5x5 sources
/// around it (staged in shared memory with a 2-px halo) in raster order of the source, so for a given (target, slot) the
/// last writer in raster order wins (oracle MODE_SNAPSHOT rule). A source at window position (k,dy) targets this pixel iff
/// its stored offset index equals (2-dy)*5 + (4-k). Phase A rewrites the whole intent plane every frame (NO_INTENT where
/// nothing was queued), so no has-intent mask is needed.
struct PhaseBArgs {
    int W, H, Wp, WW, CH;
    size_t plane;
    void* bg;
    const void* last_color;    // == this frame's colour for every pixel that queued a write
    const void* last_desc;     // == this frame's intra descriptors for every pixel that queued a write
    const ushort* intents;
    const uchar* own_slot;     // SuBSENSE: queued own-sample writes (applied before the neighbour writes, which override them); LOBSTER: nullptr
    const FrameCtl* ctl; uint32_t pending_seq; // pending_seq != 0: skip when FrameCtl::nb_applied_seq says these writes are already in the model
    uint32_t* bump_frame;      // LOBSTER: FrameCtl::frame_idx, advanced by one thread here (the frame's pixel pass is over; saves a launch)
    uint2* cbox;               // SuBSENSE: colour boxes, grown with every record written here (nullptr: not kept)
};

/// RAD: radius of the neighbourhood the intents were drawn from (2: SuBSENSE's 5x5 / 3x3 switch; 1: LOBSTER always draws from the 3x3 pattern,
/// so 9 window positions are tested per target instead of 25)
template<int CH, int RAD = 2>
__global__ void __launch_bounds__(256) neighbor_write_phaseB(const PhaseBArgs A) {
    typedef typename Pack<CH>::Col Col;
    typedef typename Pack<CH>::Desc Desc;
    typedef typename Pack<CH>::Rec Rec;
    constexpr int TW = 32 + 4, TH = 8 + 4;
    __shared__ ushort s_int[TH][TW + 2]; // +2: row pitch of 19 words (odd) keeps the 5-row column walk conflict-free
    if(A.pending_seq != 0u && A.ctl->nb_applied_seq == A.pending_seq) return;
    const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 8;
    const int tid = threadIdx.y * 32 + threadIdx.x;
    for(int i = tid; i < TW * TH; i += 256) {
        const int r = i / TW, cc = i - r * TW;
        const int gx = x0 - 2 + cc, gy = y0 - 2 + r;
        s_int[r][cc] = (gx >= 0 && gx < A.W && gy >= 0 && gy < A.H) ? A.intents[(size_t)gy * A.Wp + gx] : (ushort)NO_INTENT;
    }
    __syncthreads();
    if(A.bump_frame && tid == 0 && blockIdx.x == 0 && blockIdx.y == 0) *A.bump_frame += 1u;
    const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
    if(A.own_slot && x < A.W && y < A.H) { // the pixel's own queued update first
        const size_t p = (size_t)y * A.Wp + x;
        const uint32_t own = A.own_slot[p];
        if(own != 0xFFu) {
            const Col oc = ((const Col*)A.last_color)[p];
            ((Rec*)A.bg)[(size_t)own * A.plane + p] = rec_make(oc, ((const Desc*)A.last_desc)[p]);
            if(A.cbox) { uint2 bx = A.cbox[p]; cbox_grow(bx, col_as_u32_(oc)); A.cbox[p] = bx; }
        }
    }
    if(x < 2 || y < 2 || x > A.W - 3 || y > A.H - 3) return; // clamped targets never leave [2,dim-3]
    // pass 1: which of the 25 sources aim at this pixel (no global access); bit i = window position i = (dy+2)*5 + k
    uint32_t hits = 0;
#pragma unroll
    for(int dy = -RAD; dy <= RAD; ++dy) {
#pragma unroll
        for(int k = 2 - RAD; k <= 2 + RAD; ++k) {
            const uint32_t it = s_int[threadIdx.y + 2 + dy][threadIdx.x + k];
            if((it >> 8) == (uint32_t)((2 - dy) * 5 + (4 - k))) hits |= 1u << ((dy + 2) * 5 + k);
        }
    }
    // pass 2: apply them in raster order of the source (ascending bit index), two per round so that the source loads of both
    // are in flight together. A warp runs as many rounds as its busiest lane needs (1-2), not one per window position.
    const Col* lcol = (const Col*)A.last_color;
    const Desc* ldes = (const Desc*)A.last_desc;
    const size_t tpix = (size_t)y * A.Wp + x;
    uint2 bx = cbox_full();
    const bool keep_box = A.cbox != nullptr && hits != 0u;
    if(keep_box) bx = A.cbox[tpix];
    while(hits) {
        const int i0 = __ffs(hits) - 1;
        hits &= hits - 1;
        const int i1 = hits ? __ffs(hits) - 1 : -1;
        if(i1 >= 0) hits &= hits - 1;
        const int r0 = i0 / 5, k0 = i0 - r0 * 5;
        const uint32_t it0 = s_int[threadIdx.y + r0][threadIdx.x + k0];
        const size_t q0 = (size_t)(y + r0 - 2) * A.Wp + (x - 2 + k0);
        const Col c0 = lcol[q0]; const Desc d0 = ldes[q0];
        Col c1 = Col(); Desc d1 = Desc(); uint32_t it1 = 0;
        if(i1 >= 0) {
            const int r1 = i1 / 5, k1 = i1 - r1 * 5;
            it1 = s_int[threadIdx.y + r1][threadIdx.x + k1];
            const size_t q1 = (size_t)(y + r1 - 2) * A.Wp + (x - 2 + k1);
            c1 = lcol[q1]; d1 = ldes[q1];
        }
        const size_t dst0 = (size_t)(it0 & 0xFFu) * A.plane + tpix;
        ((Rec*)A.bg)[dst0] = rec_make(c0, d0);
        cbox_grow(bx, col_as_u32_(c0));
        if(i1 >= 0) {
            const size_t dst1 = (size_t)(it1 & 0xFFu) * A.plane + tpix;
            ((Rec*)A.bg)[dst1] = rec_make(c1, d1);
            cbox_grow(bx, col_as_u32_(c1));
        }
    }
    if(keep_box) A.cbox[tpix] = bx;
}

/// refreshModel (SuBSENSE.cpp:80-105 / LOBSTER.cpp:410-441): one thread per ROI pixel; runs only when
/// ctl->do_refresh is set (the frame tail decides on the device). Also applies the "T(x)=1" reset (:592).
struct RefreshArgs {
    int W, H, Wp, WW, CH, N;
    size_t plane;
    void* bg;
    const void* last_color; void* last_desc;
    const uint32_t* roi_bits; const uint32_t* lastfg_bits;
    float4* maps;
    const uchar* lut;
    FrameCtl* ctl;
    uint64_t seed;
    int recompute_desc;        // LOBSTER: descriptor of the sampled pixel is recomputed from last_color
    const ushort* intents; uint32_t pending_seq; // SuBSENSE: neighbour writes of this frame not yet in the model are applied first
    const uchar* own_slot;     // SuBSENSE: ... and before them the queued own-sample writes
    uint2* cbox;               // SuBSENSE: colour boxes (nullptr: not kept): rebuilt for every pixel this kernel writes to
};

template<int CH>
__global__ void __launch_bounds__(256) refresh_model_kernel(const RefreshArgs A) {
    typedef typename Pack<CH>::Col Col;
    typedef typename Pack<CH>::Desc Desc;
    typedef typename Pack<CH>::Rec Rec;
    FrameCtl* ctl = A.ctl;
    if(!ctl->do_refresh) return;
    // launched every frame (the request is decided on the device, by the frame tail): a CTA walks 32x8 tiles grid-stride. When a
    // request fires, the tail has already waited for this frame's final mask (lastfg) before the feedback kernel ended.
    const bool apply_nb = A.pending_seq != 0u && ctl->nb_applied_seq != A.pending_seq;
    const int tiles_x = A.Wp / 32, ntiles = tiles_x * ((A.H + 7) / 8);
    for(int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int x = (tile % tiles_x) * 32 + threadIdx.x, y = (tile / tiles_x) * 8 + threadIdx.y;
    if(x >= A.W || y >= A.H) continue;
    const size_t pix = (size_t)y * A.Wp + x;
    bool wrote = false;
    if(apply_nb && A.own_slot) {
        const uint32_t own = A.own_slot[pix];
        if(own != 0xFFu) { ((Rec*)A.bg)[(size_t)own * A.plane + pix] = rec_make(((const Col*)A.last_color)[pix], ((const Desc*)A.last_desc)[pix]); wrote = true; }
    }
    if(apply_nb && x >= 2 && y >= 2 && x <= A.W - 3 && y <= A.H - 3) {
        // the reference applies the frame's neighbour writes inside its pixel loop, i.e. before refreshModel: same rule as phase B
        for(int dy = -2; dy <= 2; ++dy) for(int k = 0; k < 5; ++k) {
            const size_t q = (size_t)(y + dy) * A.Wp + (x - 2 + k);
            const uint32_t it = A.intents[q];
            if((it >> 8) == (uint32_t)((2 - dy) * 5 + (4 - k))) {
                const size_t dst = (size_t)(it & 0xFFu) * A.plane + pix;
                ((Rec*)A.bg)[dst] = rec_make(((const Col*)A.last_color)[q], ((const Desc*)A.last_desc)[q]);
                wrote = true;
            }
        }
    }
    if(ctl->set_T_one && A.maps) A.maps[pix * 2].x = 1.0f;
    // a pixel that is not resampled below but received queued writes: its colour box is rebuilt from its samples (own stores are visible to own loads)
    const bool roi_px = (A.roi_bits[y * A.WW + (x >> 5)] >> (x & 31)) & 1u;
    const bool force = ctl->refresh_force != 0;
    const bool skip = !roi_px || (!force && ((__ldcg(A.lastfg_bits + y * A.WW + (x >> 5)) >> (x & 31)) & 1u));
    if(skip) { if(A.cbox && wrote) A.cbox[pix] = cbox_of_model<CH>(A.bg, A.plane, pix, A.N); continue; }
    const uint32_t N = (uint32_t)A.N, start = ctl->refresh_start, count = ctl->refresh_count, epoch = ctl->refresh_epoch;
    const uint32_t pixid = (uint32_t)(y * A.W + x);
    const Col* lc = (const Col*)A.last_color;
    uint4 rnd = make_uint4(0, 0, 0, 0);
    uint32_t have_blk = 0xFFFFFFFFu;
    for(uint32_t i = start; i < start + count; ++i) {
        const uint32_t rs = i % N;
        if((rs >> 2) != have_blk) { have_blk = rs >> 2; rnd = philox_block(A.seed, epoch, pixid, have_blk, DOM_REFRESH); }
        const uint32_t r = (rs & 3) == 0 ? rnd.x : (rs & 3) == 1 ? rnd.y : (rs & 3) == 2 ? rnd.z : rnd.w;
        int sx, sy;
        sample_pos_7x7(r, sx, sy, x, y, A.W, A.H);
        if(!force && ((__ldcg(A.lastfg_bits + sy * A.WW + (sx >> 5)) >> (sx & 31)) & 1u)) continue;
        const size_t sp = (size_t)sy * A.Wp + sx;
        const Col col = lc[sp];
        Desc d;
        if(A.recompute_desc) {
            uint32_t dd[CH];
#pragma unroll
            for(int c = 0; c < CH; ++c) {
                Lookup16 L;
#pragma unroll
                for(int q = 0; q < 4; ++q) {
                    uint32_t w = 0;
#pragma unroll
                    for(int b = 0; b < 4; ++b) {
                        const int n = q * 4 + b;
                        const Col nc = lc[(size_t)(sy + c_lbsp_dy[n]) * A.Wp + (sx + c_lbsp_dx[n])];
                        w |= col_get(nc, c) << (8 * b);
                    }
                    L.w[q] = w;
                }
                const uint32_t ref = col_get(col, c);
                dd[c] = lbsp_threshold(L, ref, A.lut[ref]);
            }
            if constexpr (CH == 1) d = (ushort)dd[0]; else d = make_uint2(dd[0] | (dd[1] << 16), dd[2]);
            ((Desc*)A.last_desc)[sp] = d; // idempotent: a pure function of last_color
        } else d = ((const Desc*)A.last_desc)[sp];
        ((Rec*)A.bg)[(size_t)rs * A.plane + pix] = rec_make(col, d);
    }
    if(A.cbox) A.cbox[pix] = cbox_of_model<CH>(A.bg, A.plane, pix, A.N);   // exact box of what the refresh leaves behind
    }
    // the last CTA to finish retires the request (every CTA has read it by then) and bumps the epoch it consumed
    __syncthreads();
    if(threadIdx.x == 0 && threadIdx.y == 0) {
        __threadfence();
        if(atomicAdd(&ctl->refresh_blocks, 1u) == gridDim.x - 1u) {
            ctl->refresh_epoch += 1; ctl->do_refresh = 0; ctl->set_T_one = 0; ctl->refresh_blocks = 0;
            if(A.pending_seq) ctl->nb_applied_seq = A.pending_seq;
        }
    }
}

/// frame-level analysis (SuBSENSE.cpp:567-583): 8x8 area mean -> two EMAs -> sum of truncated max-channel |ST-LT|
struct DownsampleArgs {
    int W, H, CH, dsW, dsH;
    const uchar* img; size_t ipitch;
    float* dsLT; float* dsST;
    FrameCtl* ctl;
};
/// one cell of OpenCV's general INTER_AREA table (imgproc resize.cpp computeResizeAreaTab) for destination index d: the source
/// range [s_first, s_first + n) and the weights of its first / inner / last entries; evaluated in double then cast, as OpenCV does
struct AreaCell { int s_first, n; float a_first, a_mid, a_last; bool has_first, has_last; };
__device__ __forceinline__ AreaCell area_cell(int d, int ssize, double scale) {
    const double fs1 = d * scale, fs2 = fs1 + scale, cell = fmin(scale, ssize - fs1);
    int s1 = (int)ceil(fs1), s2 = (int)floor(fs2);
    s2 = min(s2, ssize - 1); s1 = min(s1, s2);
    AreaCell c;
    c.has_first = (s1 - fs1) > 1e-3; c.has_last = (fs2 - s2) > 1e-3;
    c.a_first = (float)((s1 - fs1) / cell); c.a_mid = (float)(1.0 / cell); c.a_last = (float)(fmin(fmin(fs2 - s2, 1.), cell) / cell);
    c.s_first = c.has_first ? s1 - 1 : s1;
    c.n = (s2 - s1) + (c.has_first ? 1 : 0) + (c.has_last ? 1 : 0);
    return c;
}
__device__ __forceinline__ float area_weight(const AreaCell& c, int i) { // weight of the i-th table entry of the cell
    if(c.has_first && i == 0) return c.a_first;
    if(c.has_last && i == c.n - 1) return c.a_last;
    return c.a_mid;
}
/// cv::resize(INTER_AREA) of one destination pixel for a non-integer shrink factor, with OpenCV's float accumulation order
/// (ResizeArea_Invoker<uchar,float>: row buffer over the column table, then beta-weighted row sums)
template<int CH>
__device__ __forceinline__ void area_general_pixel(const uchar* img, size_t ipitch, int W, int H, int dsW, int dsH, int dx, int dy, float (&v)[CH]) {
    const double scale_x = 1. / ((double)dsW / W), scale_y = 1. / ((double)dsH / H);
    const AreaCell cx = area_cell(dx, W, scale_x), cy = area_cell(dy, H, scale_y);
    float sum[CH];
    for(int j = 0; j < cy.n; ++j) {
        const uchar* row = img + (size_t)(cy.s_first + j) * ipitch;
        float buf[CH];
#pragma unroll
        for(int c = 0; c < CH; ++c) buf[c] = 0.f;
        for(int i = 0; i < cx.n; ++i) {
            const float a = area_weight(cx, i);
#pragma unroll
            for(int c = 0; c < CH; ++c) buf[c] = __fadd_rn(buf[c], __fmul_rn((float)row[(size_t)(cx.s_first + i) * CH + c], a));
        }
        const float b = area_weight(cy, j);
#pragma unroll
        for(int c = 0; c < CH; ++c) sum[c] = j == 0 ? __fmul_rn(b, buf[c]) : __fadd_rn(sum[c], __fmul_rn(b, buf[c]));
    }
#pragma unroll
    for(int c = 0; c < CH; ++c) v[c] = fminf(fmaxf(rintf(sum[c]), 0.f), 255.f);
}

template<int CH>
__global__ void __launch_bounds__(128) downsample_motion_kernel(const DownsampleArgs A) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long diff = 0;
    if(i < A.dsW * A.dsH && A.ctl->lr_scaling) {
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
        uint32_t best = 0;
#pragma unroll
        for(int c = 0; c < CH; ++c) {
            const float v = vv[c];
            const size_t k = (size_t)i * CH + c;
            const float lt = __fadd_rn(__fmul_rn(v, aLT), __fmul_rn(A.dsLT[k], __fsub_rn(1.0f, aLT)));
            const float st = __fadd_rn(__fmul_rn(v, aST), __fmul_rn(A.dsST[k], __fsub_rn(1.0f, aST)));
            A.dsLT[k] = lt; A.dsST[k] = st;
            uint32_t d = (uint32_t)fabsf(__fsub_rn(st, lt));
            if(CH == 1) d >>= 1;
            best = max(best, d);
        }
        diff = best;
    }
#pragma unroll
    for(int o = 16; o > 0; o >>= 1) diff += __shfl_xor_sync(0xFFFFFFFFu, diff, o);
    if((threadIdx.x & 31) == 0 && diff) atomicAdd(&A.ctl->tot_color_diff, diff);
}

} // namespace lvb
