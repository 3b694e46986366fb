// litiv_b200 — device-resident state of one video stream (one background subtractor instance).
//
// HBM layout (Wp = W rounded up to 32 pixels so every row starts on a 32-bit mask word; WW = Wp/32):
//   model colour  : [N][H][Wp]  3ch: u32 (B,G,R,0) / 1ch: u8         sample-major SoA, a warp reads 128 B per sample row
//   model desc    : [N][H][Wp]  3ch: uint2 (d0|d1<<16, d2) / 1ch: u16
//   feedback maps : [H][Wp] x 8 f32 interleaved (T,R,v,Dlast | DminLT,DminST,rawLT,rawST) = 2 x float4 per pixel
//   final-segm EMA: [H][Wp] float2 (LT,ST)
//   last colour / last desc: [H][Wp] same packing as one model sample
//   masks         : bit-packed [H][WW] u32 (roi, raw, lastraw, lastrawblink, blinks, lastfg, dilinv, unstable, ghost x2, intent bits)
//   nb intents    : [H][Wp] u16 (neighbour code << 8 | slot), valid where the intent bit is set
//   frame scalars : FrameCtl (device) so that a whole frame needs no host round trip
#pragma once
#include "common.cuh"

namespace lvb {

struct Params { // mirrors include/litiv_b200.h : lvb_params
    float rel_lbsp_threshold;
    int lbsp_threshold_offset;
    int desc_dist_threshold;
    int color_dist_threshold;
    int n_samples;
    int n_required;
    int n_samples_for_moving_avgs;
    int n_global_words;
    int median_blur_kernel_size;
};

struct FrameCtl {
    uint32_t frame_idx;        // index of the frame being processed (1-based, == reference m_nFrameIdx after ++)
    float aLT, aST;            // rolling average factors of the current frame
    float t_lower, t_upper;    // current learning-rate caps
    uint32_t cooldown, frames_since_reset, auto_reset, lr_scaling, use3x3;
    int32_t median_k;
    float last_nonzero_ratio;
    uint32_t nonzero_count;    // accumulators, cleared by the tail kernel
    unsigned long long tot_color_diff;
    uint32_t do_refresh, refresh_epoch, refresh_start, refresh_count, refresh_force, set_T_one;
    uint32_t flood_changed[4];   // 3 rotating convergence flags of pp_flood (+1 pad)
    uint32_t roi_count;
    unsigned long long stat_scanned, stat_writes, stat_fg; // optional instrumentation
    uint32_t blocks_done;      // ticket counter of the feedback kernel (its last CTA runs the frame tail)
    uint32_t refresh_blocks;   // ticket counter of the conditional refresh kernel (its last CTA retires the request)
    uint32_t chain_done;       // sequence number of the last frame whose final mask (lastfg) is complete; written on the mask stream
    uint32_t nb_applied_seq;   // sequence number of the last frame whose queued neighbour writes are already in the model
    uint32_t wl_count;         // SuBSENSE scan work-list: pixels still undecided after the two prefetched samples (reset by the frame tail)
    uint32_t wl2_count;        // ... and those still undecided after the first tail pass
    uint32_t wl_cursor, wl2_cursor;   // chunk cursors of the two tail passes (warps pull 32 entries at a time)
    uint32_t spin_timeout;     // != 0: sequence number of a frame whose model reset gave up waiting for its final mask (see subsense_tail_warp)
    uint32_t pad2[3];
};

// One background sample = one naturally aligned record (colour + descriptors): 16 bytes for 3 channels, 4 bytes for 1.
// A stochastic sample update is then ONE partial-sector write instead of two (scattered 32-byte read-modify-writes are what
// the update kernels pay for in DRAM), and a lone sample fetched by the scan tail is one sector instead of two.
template<int CH> struct Pack;
template<> struct Pack<3> { typedef uint32_t Col; typedef uint2 Desc; typedef uint4 Rec; };
template<> struct Pack<1> { typedef uchar Col; typedef ushort Desc; typedef uint32_t Rec; };
__device__ __forceinline__ uint4 rec_make(uint32_t c, uint2 d) { return make_uint4(c, d.x, d.y, 0u); }
__device__ __forceinline__ uint32_t rec_make(uchar c, ushort d) { return (uint32_t)c | ((uint32_t)d << 16); }
__device__ __forceinline__ uint32_t rec_col(const uint4& r) { return r.x; }
__device__ __forceinline__ uint2 rec_desc(const uint4& r) { return make_uint2(r.y, r.z); }
__device__ __forceinline__ uchar rec_col(const uint32_t& r) { return (uchar)(r & 0xFFu); }
__device__ __forceinline__ ushort rec_desc(const uint32_t& r) { return (ushort)(r >> 16); }

__device__ __forceinline__ uint32_t desc_get(const uint2& d, int c) { return c == 0 ? (d.x & 0xFFFFu) : c == 1 ? (d.x >> 16) : (d.y & 0xFFFFu); }
__device__ __forceinline__ uint32_t desc_get(const ushort& d, int) { return d; }
__device__ __forceinline__ uint32_t col_get(const uint32_t& v, int c) { return (v >> (8 * c)) & 0xFFu; }
__device__ __forceinline__ uint32_t col_get(const uchar& v, int) { return v; }
__device__ __forceinline__ uint32_t col_as_u32_(const uint32_t& v) { return v; }
__device__ __forceinline__ uint32_t col_as_u32_(const uchar& v) { return v; }

struct SubArgs {
    int W, H, Wp, WW, N, REQ;
    size_t plane;              // H*Wp
    const uchar* img; size_t ipitch;  // current frame, interleaved bytes
    void* bg;                  // sample records [N][H][Wp]
    float4* maps;              // 2 float4 per pixel
    float* r_plane;            // compact copy of R(x) for the scan kernel (which would otherwise pull a 32-byte sector per pixel for 4 bytes)
    float2* fin;               // final-segmentation EMAs (SuBSENSE.cpp:553-554). The feedback kernel of frame k+1 folds frame k's
    uint32_t ema_frame;        // final mask into them before it uses them (ema_frame = k, 0: nothing pending) and stores them back
    int avg_samples;
    uint2* hand;               // scan -> feedback hand-off word
    void* last_color; void* last_desc;          // this frame's colour / intra descriptors (written by the scan)
    const void* prev_color; const void* prev_desc; // previous frame's (read by the scan: D_last and the pending neighbour writes)
    uint32_t pending_seq;      // frame whose queued neighbour writes (intents[]) the scan has to apply first (0: none)
    const uint32_t* roi_bits;
    uint32_t* raw_bits; uint32_t* unstable_bits; const uint32_t* blinks_bits; const uint32_t* lastfg_bits;
    const uint32_t* ghost_prev; uint32_t* ghost_cur;
    uint32_t* intent_bits; ushort* intents; size_t bitplane;   // 5 intent planes (one per row offset) of `bitplane` words
    const uchar* lut;
    FrameCtl* ctl;
    uint64_t seed;
    uint32_t lr_fixed;         // 0: use ceil(T(x)) ; else the ceil'd override (0xFFFFFFFF for +inf)
    int min_color, desc_off;
    int use_tma, collect_stats;
    uint32_t n_magic;          // floor(2^32 / N) for fast_mod
    uint2* cbox;               // SuBSENSE: per-pixel colour bounding box of the sample model (ColorBox, subsense.cuh); nullptr: not kept
    const uint32_t* magic;     // [257] floor(2^32 / n) (n = 1: 0xFFFFFFFF), device table for x % ceil(T(x))
    const float* div_color;    // [colorRange + 1] i / colorRange  (IEEE quotients tabulated on the host: the feedback step
    const float* div_desc;     // [descRange + 1]  i / descRange    normalises four small integers per pixel)
    uint32_t lr_magic, lr2_magic; // magic numbers of lr_fixed and lr_fixed/2+1 when the rate is fixed
    // scan work-list (see subsense.cuh): contexts of the pixels the scan kernel could not decide with the two prefetched samples,
    // structure-of-arrays [WlCtx<CH>::FIELDS][wl_cap] u32, and the indices of the entries that survive the first tail pass
    uint32_t* wl_ctx; uint32_t wl_cap; uint32_t* wl2_idx;
    uchar* own_slot;           // [H][Wp] own-sample write queued by the feedback kernel (slot, 0xFF: none); applied by the next scan
};

} // namespace lvb
