// litiv_b200 — host side of the C ABI (include/litiv_b200.h): per-stream context, HBM layout, kernel sequencing.
// Single translation unit: the kernels live in the .cuh files included below.
#include "../../include/litiv_b200.h"
#include "lobster.cuh"
#include "postproc.cuh"
#include "pawcs.cuh"
#include "metrics.cuh"
#include <string>
#include <vector>
#include <stdexcept>
#include <cmath>
#include <cstring>
#include <atomic>
#include <algorithm>
#include <cstdlib>
#include <cstdio>
#include <thread>
#include <mutex>
#include <condition_variable>
#include <functional>

using namespace lvb;

namespace {

thread_local std::string g_err;
std::atomic<uint64_t> g_launches{0};

#define CK(call) do { cudaError_t e_ = (call); if(e_ != cudaSuccess) throw std::runtime_error(std::string("CUDA error: ") + cudaGetErrorString(e_) + " at " #call); } while(0)
#define LAUNCHED() do { ++g_launches; CK(cudaGetLastError()); } while(0)
#define REQUIRE(cond, msg) do { if(!(cond)) throw std::runtime_error(msg); } while(0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if(!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = (EncodeTiledFn)p;
        (void)cudaGetLastError();
    }
    return fn;
}
/// tensor map over an interleaved byte image so one TMA box = (TILE_W+4)x(TILE_H+4) pixels incl. the LBSP halo
bool make_image_tmap(CUtensorMap* map, const void* ptr, int W, int H, int C, size_t pitch) {
    std::memset(map, 0, sizeof(*map));
    static const bool disabled = getenv("LVB_NO_TMA") != nullptr; // debugging aid: force the cooperative-copy staging path
    EncodeTiledFn fn = disabled ? nullptr : get_encode_fn();
    if(!fn || ((uintptr_t)ptr & 15) || (pitch & 15)) return false;
    const cuuint64_t gdim[2] = {(cuuint64_t)W * C, (cuuint64_t)H};
    const cuuint64_t gstr[1] = {(cuuint64_t)pitch};
    const cuuint32_t box[2] = {(cuuint32_t)tile_pitch(C), (cuuint32_t)TILE_ROWS};
    const cuuint32_t estr[2] = {1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

/// tensor map over a per-pixel plane [rows][width] of `esz`-byte elements (esz 1, 2, 4 or 8), box = box_w x box_h elements
bool make_plane_tmap(CUtensorMap* map, const void* ptr, int esz, size_t width, size_t rows, int box_w, int box_h) {
    std::memset(map, 0, sizeof(*map));
    EncodeTiledFn fn = get_encode_fn();
    if(!fn || ((uintptr_t)ptr & 15) || ((width * esz) & 15) || ((size_t)box_w * esz & 15) || box_w > 256 || box_h > 256) return false;
    const CUtensorMapDataType dt = esz == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : esz == 2 ? CU_TENSOR_MAP_DATA_TYPE_UINT16 : esz == 4 ? CU_TENSOR_MAP_DATA_TYPE_UINT32 : CU_TENSOR_MAP_DATA_TYPE_UINT64;
    const cuuint64_t gdim[2] = {(cuuint64_t)width, (cuuint64_t)rows};
    const cuuint64_t gstr[1] = {(cuuint64_t)(width * esz)};
    const cuuint32_t box[2] = {(cuuint32_t)box_w, (cuuint32_t)box_h};
    const cuuint32_t estr[2] = {1, 1};
    return fn(map, dt, 2, const_cast<void*>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// All device work of an instance goes to ITS stream (created non-blocking): the legacy default stream is never used,
// because a pageable cudaMemcpy may return before its DMA has landed and a non-blocking stream would not wait for it.
template<typename T> T* dalloc(cudaStream_t st, size_t n, bool zero = true) {
    T* p = nullptr;
    CK(cudaMalloc((void**)&p, std::max<size_t>(n, 1) * sizeof(T)));
    if(zero) CK(cudaMemsetAsync(p, 0, std::max<size_t>(n, 1) * sizeof(T), st));
    return p;
}
void h2d(cudaStream_t st, void* dst, const void* src, size_t bytes) {
    CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));
}
void d2h(cudaStream_t st, void* dst, const void* src, size_t bytes) {
    CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
}

void morph_rect_host(const uint8_t* src, uint8_t* dst, int W, int H, int r) { // dilate only (ROI border ring, init-time)
    for(int y = 0; y < H; ++y) for(int x = 0; x < W; ++x) {
        uint8_t v = 0;
        for(int dy = -r; dy <= r && !v; ++dy) for(int dx = -r; dx <= r; ++dx) {
            const int yy = y + dy, xx = x + dx;
            if(yy >= 0 && yy < H && xx >= 0 && xx < W && src[(size_t)yy * W + xx]) { v = 255; break; }
        }
        dst[(size_t)y * W + x] = v;
    }
}

} // namespace

struct lvb_context {
    int algo = 0, device = 0;
    Params P{};
    uint64_t seed = 0;
    int W = 0, H = 0, C = 0, Wp = 0, WW = 0, dsW = 0, dsH = 0;
    size_t plane = 0;
    cudaStream_t stream = nullptr;
    bool initialized = false;
    int uf_rs = 0; uint32_t* uf_parent = nullptr; ushort* uf_rankbase = nullptr;
    // frame staging
    uint8_t* d_img = nullptr; size_t ipitch = 0; CUtensorMap tmap_img; int use_tma = 0;
    const uint8_t* ext_ptr = nullptr; size_t ext_pitch = 0; CUtensorMap tmap_ext; int ext_tma = 0;
    uint8_t* d_mask = nullptr;
    uint8_t* h_img = nullptr; uint8_t* h_mask = nullptr; uint8_t* user_mask = nullptr;
    // model + maps
    void* bg = nullptr;       // sample records [N][H][Wp] (Pack<C>::Rec: 16 bytes for 3 channels, 4 for 1)
    float4* maps = nullptr; float2* fin = nullptr;
    void* last_color = nullptr; void* last_desc = nullptr; void* tmp_desc = nullptr;
    // SuBSENSE: the scan of frame k+1 applies the neighbour writes queued by frame k, whose sources are frame k's colours and
    // descriptors, while it writes frame k+1's: last_color/last_desc name the LATEST frame's planes, *_alt the spare ones
    void* last_color_alt = nullptr; void* last_desc_alt = nullptr;
    uint32_t nb_seq = 0;                // sequence number of the frame whose queued neighbour writes may still be pending (0: none)
    uint32_t fin_pending = 0;           // SuBSENSE: frame index whose final mask is not folded into the final-segmentation EMAs yet (0: none)
    uint32_t* bits = nullptr; // all bit planes, one allocation
    uint32_t *roi_bits, *raw, *lastraw, *lastrawblink, *blinks, *tmpA, *pre, *reach, *comb, *lastfg, *dilinv, *unstable, *ghost[2], *intent_bits;
    int ghost_idx = 0;
    // SuBSENSE runs the mask chain of frame k beside the feedback kernel of frame k and the scan of frame k+1: what the chain
    // writes and the feedback kernel reads (previous frame's value) is double-buffered. raw/blinks/lastfg/fin always name the
    // buffers of the LATEST frame, *_alt the spare ones the next frame will write.
    uint32_t *raw_alt = nullptr, *blinks_alt = nullptr, *lastfg_alt = nullptr; float2* fin_alt = nullptr;
    uint2* hand = nullptr;              // scan -> feedback hand-off
    cudaStream_t s_post = nullptr;      // mask chain (high priority: its small kernels slip in beside the scan of the next frame)
    cudaEvent_t ev_scan = nullptr, ev_post = nullptr, ev_ds = nullptr, ev_mask = nullptr;
    bool post_pending = false;          // ev_post was recorded by a frame whose chain the next feedback kernel has to wait for
    uint32_t sub_frame = 1;             // host mirror of FrameCtl::frame_idx (SuBSENSE; the final-mask EMA factors derive from it)
    uint32_t chain_seq = 0;             // frames enqueued (sequence number written to FrameCtl::chain_done by the mask stream)
    ushort* intents = nullptr;
    ScanMaps scan_maps[2];              // SuBSENSE scan kernel: TMA maps of its per-pixel planes; [i] reads colour / descriptor plane pair i as "previous frame"
    int scan_maps_idx = 0;              // which pair is the latest frame's (follows the last_color / last_color_alt swap)
    uint8_t* own_slot = nullptr;        // SuBSENSE: queued own-sample writes (slot per pixel, 0xFF none), applied by the next scan
    uint2* cbox = nullptr;              // SuBSENSE: per-pixel colour bounding box of the samples (subsense.cuh: ColorBox); LVB_NO_CBOX=1 disables it
    uint32_t* wl_ctx = nullptr; uint32_t* wl2_idx = nullptr; uint32_t wl_cap = 0;   // SuBSENSE scan work-list (subsense.cuh: WlCtx)
    uint8_t* lut = nullptr;
    bool lut_small = false;   // every LUT entry (now and after any +-1 adaptation) is <= 127: the kernels take the 7-bit compare path
    uint32_t* magic = nullptr; // [257] floor(2^32 / n)
    float* r_plane = nullptr;  // compact R(x) plane read by the SuBSENSE scan kernel
    uint8_t* eval_gt = nullptr; uint8_t* eval_roi = nullptr; unsigned long long* eval_cnt = nullptr; // lvb_binclassif_accumulate scratch (W*H each)
    float* div_tab = nullptr;  // [colorRange + 1] i / colorRange, then [descRange + 1] i / descRange
    FrameCtl* ctl = nullptr;
    float* dsLT = nullptr; float* dsST = nullptr;
    std::vector<uint8_t> roi_host;
    size_t orig_roi_count = 0, roi_count = 0;
    int collect_stats = 0, median_k = 9;
    int sm_count = 148;       // persistent kernels size their grids as sm_count x resident CTAs per SM
    uint64_t stat_frames = 0;
    bool pending = false;
    // two-deep host pipeline (lvb_apply_async): slot k%2 = {device frame, device mask, pinned staging}; uploads on s_in, masks back on s_out
    struct Slot { uint8_t* d_img = nullptr; uint8_t* d_mask = nullptr; uint8_t* h_img = nullptr; uint8_t* h_mask = nullptr; CUtensorMap tmap; int use_tma = 0;
                  cudaEvent_t h2d_done = nullptr, compute_done = nullptr, d2h_done = nullptr, in_consumed = nullptr; uint8_t* user_mask = nullptr; bool direct = false, busy = false, used = false; };
    Slot slot[2];
    cudaStream_t s_in = nullptr, s_out = nullptr;
    cudaStream_t s_aux = nullptr; cudaEvent_t ev_fork = nullptr, ev_join = nullptr; // phase B runs beside the mask post-processing
    uint64_t n_submitted = 0, n_collected = 0;
    bool profile = false; std::vector<cudaEvent_t> prof_events; double prof_ms = 0; uint64_t prof_n = 0;
    std::vector<cudaEvent_t> prof2_events; double prof2_ms = 0; uint64_t prof2_n = 0; // second-largest kernel (SuBSENSE feedback)
    std::vector<cudaEvent_t> prof3_events; double prof3_ms = 0; uint64_t prof3_n = 0; // SuBSENSE scan tail passes
    bool direct_mask = false;
    // LVB_TRACE=1 (debugging aid): an event after every kernel of the main stream; lvb_get_profile prints the per-segment averages
    bool trace_on = getenv("LVB_TRACE") != nullptr; std::vector<std::pair<const char*, cudaEvent_t>> trace;
    // PAWCS
    int NW = 0, NG = 0, gW = 0, gH = 0;
    uint32_t paw_frame = 1;   // host mirror of FrameCtl::frame_idx (decides which frames run maintenance / the 500-frame check)
    uint2* lw_key = nullptr; void* lw_rec = nullptr;   // key = (occurrences, first + last), record = (colour, descriptors, first): pawcs.cuh
    uint8_t* glut = nullptr; float *gmap = nullptr, *gmap_tmp = nullptr; GDict* gd = nullptr;
    uint32_t *roi255 = nullptr, *illum = nullptr, *did = nullptr, *dil = nullptr, *gop_bits = nullptr;
    uint4* paw_intents = nullptr; float* gop_w = nullptr; uint8_t* gop_g = nullptr; uint8_t* ds_roi = nullptr; uint8_t* bgimg = nullptr;
    size_t ds_roi_count = 0;

    size_t col_bytes() const { return C == 1 ? 1 : 4; }
    size_t desc_bytes() const { return C == 1 ? 2 : 8; }
    size_t rec_bytes() const { return C == 1 ? 4 : 16; }
    size_t paw_rec_bytes() const { return C == 1 ? 8 : 16; }   // PawRec<C>::T

    void free_all() {
        void* ptrs[] = {cbox, own_slot, wl_ctx, wl2_idx, eval_gt, eval_roi, eval_cnt, r_plane, div_tab, last_color_alt, last_desc_alt, hand, fin_alt, magic, uf_parent, uf_rankbase, d_img, d_mask, bg, maps, fin, last_color, last_desc, tmp_desc, bits, intents, lut, ctl, dsLT, dsST,
                        lw_rec, lw_key, glut, gmap, gmap_tmp, gd, paw_intents, gop_w, gop_g, ds_roi, bgimg};
        for(void* p : ptrs) if(p) cudaFree(p);
        d_img = nullptr; d_mask = nullptr; bg = nullptr; maps = nullptr; fin = nullptr; last_color = last_desc = tmp_desc = nullptr;
        own_slot = nullptr; cbox = nullptr; wl_ctx = nullptr; wl2_idx = nullptr; wl_cap = 0;
        eval_gt = eval_roi = nullptr; eval_cnt = nullptr; r_plane = nullptr; div_tab = nullptr; last_color_alt = last_desc_alt = nullptr; nb_seq = 0; fin_pending = 0; hand = nullptr; fin_alt = nullptr; post_pending = false; magic = nullptr; uf_parent = nullptr; uf_rankbase = nullptr; bits = nullptr; intents = nullptr; lut = nullptr; ctl = nullptr; dsLT = dsST = nullptr;
        lw_rec = nullptr; lw_key = nullptr; glut = nullptr; gmap = gmap_tmp = nullptr; gd = nullptr;
        paw_intents = nullptr; gop_w = nullptr; gop_g = nullptr; ds_roi = nullptr; bgimg = nullptr;
        if(h_img) cudaFreeHost(h_img);
        if(h_mask) cudaFreeHost(h_mask);
        h_img = h_mask = nullptr;
        if(slot[1].d_img) cudaFree(slot[1].d_img);
        if(slot[1].d_mask) cudaFree(slot[1].d_mask);
        if(slot[1].h_img) cudaFreeHost(slot[1].h_img);
        if(slot[1].h_mask) cudaFreeHost(slot[1].h_mask);
        for(Slot& sl : slot) { sl.d_img = sl.d_mask = sl.h_img = sl.h_mask = nullptr; sl.busy = false; }
        n_submitted = n_collected = 0;
        initialized = false;
    }
};

namespace {

dim3 tile_grid(const lvb_context* c) { return dim3(c->Wp / 32, (c->H + 7) / 8); }
// grid / block of the kernels that stage an input tile (TILE_W x TILE_H pixels + halo) through TMA
dim3 stage_grid(const lvb_context* c) { return dim3(c->Wp / TILE_W, (c->H + TILE_H - 1) / TILE_H); }
const dim3 stage_block(TILE_W, TILE_H);
dim3 word_grid(const lvb_context* c) { return dim3((c->WW + 255) / 256, c->H); }

/// wait for every stream of the instance (the mask chain and the auxiliary kernels run beside the instance stream)
void sync_streams(lvb_context* c) {
    CK(cudaStreamSynchronize(c->stream));
    if(c->s_post) CK(cudaStreamSynchronize(c->s_post));
    if(c->s_aux) CK(cudaStreamSynchronize(c->s_aux));
}

void get_ctl(lvb_context* c, FrameCtl& f) {
    sync_streams(c);
    CK(cudaMemcpyAsync(&f, c->ctl, sizeof(FrameCtl), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
}
void put_ctl(lvb_context* c, const FrameCtl& f) {
    sync_streams(c);
    CK(cudaMemcpyAsync(c->ctl, &f, sizeof(FrameCtl), cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
}

/// true when no LUT entry can exceed 127: the current values and the cap of the +-1 adaptation (SuBSENSE.cpp:563, PAWCS.cpp:1470)
bool lut_fits_7bit(const lvb_context* c, const uint8_t* lut) {
    for(int i = 0; i < 256; ++i) if(lut[i] > 127) return false;
    const float hi = std::fmin(std::fmax(std::rint((float)c->P.lbsp_threshold_offset + 255.0f * c->P.rel_lbsp_threshold), 0.f), 255.f);
    return hi <= 127.f;
}
uint32_t magic_of(uint32_t n) { return n <= 1u ? 0xFFFFFFFFu : (uint32_t)(0x100000000ull / n); }

__global__ void mark_nb_applied_kernel(FrameCtl* ctl, uint32_t seq) { ctl->nb_applied_seq = seq; }
/// SuBSENSE leaves the neighbour writes of the latest frame queued for the next frame's scan; anything else that reads the sample
/// model (state export, getBackgroundImage, a host-requested refreshModel) applies them first with the standalone kernel
void flush_pending(lvb_context* c) {
    if(c->algo != LVB_ALGO_SUBSENSE || !c->initialized) return;
    if(c->fin_pending) { // final-segmentation EMAs of the latest frame
        PostArgs P{};
        P.W = c->W; P.H = c->H; P.WW = c->WW; P.Wp = c->Wp; P.lastfg = c->lastfg; P.fin = c->fin; P.fin_in = c->fin; P.ctl = c->ctl;
        P.frame = c->fin_pending; P.avg_samples = c->P.n_samples_for_moving_avgs;
        pp_final_ema<<<dim3(c->Wp / 32, (c->H + 7) / 8), dim3(32, 8), 0, c->stream>>>(P); LAUNCHED();
        CK(cudaStreamSynchronize(c->stream));
        c->fin_pending = 0;
    }
    if(c->nb_seq == 0) return;
    PhaseBArgs B{};
    B.W = c->W; B.H = c->H; B.Wp = c->Wp; B.WW = c->WW; B.CH = c->C; B.plane = c->plane;
    B.bg = c->bg; B.last_color = c->last_color; B.last_desc = c->last_desc; B.intents = c->intents; B.own_slot = c->own_slot; B.cbox = c->cbox;
    B.ctl = c->ctl; B.pending_seq = c->nb_seq;
    const dim3 tg(c->Wp / 32, (c->H + 7) / 8), tb(32, 8);
    if(c->C == 1) neighbor_write_phaseB<1><<<tg, tb, 0, c->stream>>>(B); else neighbor_write_phaseB<3><<<tg, tb, 0, c->stream>>>(B);
    LAUNCHED();
    mark_nb_applied_kernel<<<1, 1, 0, c->stream>>>(c->ctl, c->nb_seq); LAUNCHED();
    CK(cudaStreamSynchronize(c->stream));
}
/// exact colour boxes from the sample model as it stands (state import, and periodically: the boxes only grow in between)
void rebuild_cbox(lvb_context* c) {
    if(!c->cbox || c->algo != LVB_ALGO_SUBSENSE) return;
    const dim3 tg(c->Wp / 32, (c->H + 7) / 8), tb(32, 8);
    if(c->C == 1) cbox_rebuild_kernel<1><<<tg, tb, 0, c->stream>>>(c->bg, c->plane, c->W, c->H, c->Wp, c->P.n_samples, c->cbox);
    else cbox_rebuild_kernel<3><<<tg, tb, 0, c->stream>>>(c->bg, c->plane, c->W, c->H, c->Wp, c->P.n_samples, c->cbox);
    LAUNCHED();
}
/// conditional refreshModel on the instance stream (exits at once unless FrameCtl::do_refresh is set)
void launch_refresh(lvb_context* c) {
    RefreshArgs R{};
    R.W = c->W; R.H = c->H; R.Wp = c->Wp; R.WW = c->WW; R.CH = c->C; R.N = c->P.n_samples; R.plane = c->plane;
    R.bg = c->bg; R.last_color = c->last_color; R.last_desc = c->last_desc;
    R.roi_bits = c->roi_bits; R.lastfg_bits = c->lastfg; R.maps = c->algo == LVB_ALGO_SUBSENSE ? c->maps : nullptr;
    R.lut = c->lut; R.ctl = c->ctl; R.seed = c->seed; R.recompute_desc = c->algo == LVB_ALGO_LOBSTER;
    const dim3 tgd = tile_grid(c);
    R.intents = c->intents; R.pending_seq = c->algo == LVB_ALGO_SUBSENSE ? c->nb_seq : 0u;
    R.own_slot = c->algo == LVB_ALGO_SUBSENSE ? c->own_slot : nullptr;
    R.cbox = c->algo == LVB_ALGO_SUBSENSE ? c->cbox : nullptr;
    const int rgrid = (int)std::min<size_t>((size_t)tgd.x * tgd.y, 148 * 8);
    if(c->C == 1) refresh_model_kernel<1><<<rgrid, dim3(32, 8), 0, c->stream>>>(R);
    else refresh_model_kernel<3><<<rgrid, dim3(32, 8), 0, c->stream>>>(R);
    LAUNCHED();
}



} // namespace

/// mask stream, right behind the median kernel: the final mask (lastfg) of frame `seq` is complete
__global__ void chain_done_kernel(FrameCtl* ctl, uint32_t seq) { ctl->chain_done = seq; }
__global__ void refresh_start_kernel(FrameCtl* ctl, uint64_t seed, uint32_t N) {
    ctl->refresh_start = philox_draw(seed, ctl->refresh_epoch, 0, 0, DOM_REFRESH_START) % N;
}

namespace {


// ---- PAWCS host side ----
PawArgs paw_args(lvb_context* c, const uint8_t* img, size_t pitch, int use_tma, double lr);
uint32_t lr_to_fixed(double lr);

void paw_launch_refresh(lvb_context* c, const PawArgs& A, uint32_t frame_off) {
    const dim3 tg = tile_grid(c), tb(32, 8);
    if(c->C == 1) pawcs_refresh_local<1><<<tg, tb, 0, c->stream>>>(A, frame_off); else pawcs_refresh_local<3><<<tg, tb, 0, c->stream>>>(A, frame_off);
    LAUNCHED();
    if(c->C == 1) pawcs_refresh_global<1><<<1, 1024, 0, c->stream>>>(A, frame_off); else pawcs_refresh_global<3><<<1, 1024, 0, c->stream>>>(A, frame_off);
    LAUNCHED();
    pawcs_glut_bubble<<<tg, tb, 0, c->stream>>>(A, 1); LAUNCHED(); // its last CTA closes the request (epoch, refresh_req)
}

/// cv::resize(INTER_AREA) of a one-channel 8-bit image to (dw, dh), dw = W/8, dh = H/8: the exact 8x8 mean when both dimensions
/// divide by 8 (OpenCV's integer-scale fast path), OpenCV's general area path otherwise (column / row tables of
/// computeResizeAreaTab, float accumulation in ResizeArea_Invoker's order; same arithmetic as csrc/subsense.cuh area_general_pixel)
void host_resize_area(const uint8_t* src, int W, int H, int dw, int dh, uint8_t* dst) {
    auto sat = [](float v) { const long q = std::lrint((double)v); return (uint8_t)(q < 0 ? 0 : q > 255 ? 255 : q); };
    if(W % 8 == 0 && H % 8 == 0) {
        for(int y = 0; y < dh; ++y) for(int x = 0; x < dw; ++x) {
            int sum = 0;
            for(int dy = 0; dy < 8; ++dy) for(int dx = 0; dx < 8; ++dx) sum += src[(size_t)(y * 8 + dy) * W + x * 8 + dx];
            dst[(size_t)y * dw + x] = sat((float)sum * (1.0f / 64));
        }
        return;
    }
    struct Cell { int first, n; float a_first, a_mid, a_last; bool hf, hl; };
    auto cell = [](int d, int ssize, double scale) {
        const double fs1 = d * scale, fs2 = fs1 + scale, cw = std::min(scale, ssize - fs1);
        int s1 = (int)std::ceil(fs1), s2 = (int)std::floor(fs2);
        s2 = std::min(s2, ssize - 1); s1 = std::min(s1, s2);
        Cell c;
        c.hf = (s1 - fs1) > 1e-3; c.hl = (fs2 - s2) > 1e-3;
        c.a_first = (float)((s1 - fs1) / cw); c.a_mid = (float)(1.0 / cw); c.a_last = (float)(std::min(std::min(fs2 - s2, 1.), cw) / cw);
        c.first = c.hf ? s1 - 1 : s1; c.n = (s2 - s1) + (c.hf ? 1 : 0) + (c.hl ? 1 : 0);
        return c;
    };
    auto weight = [](const Cell& c, int i) { return (c.hf && i == 0) ? c.a_first : (c.hl && i == c.n - 1) ? c.a_last : c.a_mid; };
    const double sx = 1. / ((double)dw / W), sy = 1. / ((double)dh / H);
    for(int y = 0; y < dh; ++y) {
        const Cell cy = cell(y, H, sy);
        for(int x = 0; x < dw; ++x) {
            const Cell cx = cell(x, W, sx);
            float sum = 0.f;
            for(int j = 0; j < cy.n; ++j) {
                float buf = 0.f;
                for(int i = 0; i < cx.n; ++i) { const float t = (float)src[(size_t)(cy.first + j) * W + cx.first + i] * weight(cx, i); buf = buf + t; }
                const float t = weight(cy, j) * buf;
                sum = j == 0 ? t : sum + t;
            }
            dst[(size_t)y * dw + x] = sat(sum);
        }
    }
}

/// PAWCS part of initialize (PAWCS.cpp:431-557); the common part (ROI, LUT, last colour / descriptor frames) is already done
void paw_initialize(lvb_context* c, FrameCtl& f, size_t orig) {
    const int W = c->W, H = c->H, C = c->C;
    REQUIRE(c->P.n_samples / 2 > 0, "max local/global word counts must be positive");
    const size_t npx = (size_t)W * H, bp = (size_t)H * c->WW;
    c->dsW = W / 8; c->dsH = H / 8; c->gW = W / 2; c->gH = H / 2;
    std::vector<uint8_t> dsr((size_t)c->dsW * c->dsH);
    host_resize_area(c->roi_host.data(), W, H, c->dsW, c->dsH, dsr.data()); // cv::resize(ROI, INTER_AREA, 1/8) (:446)
    const int maxG = c->P.n_samples / 2, qvga = 320 * 240, defk = c->P.median_blur_kernel_size;
    c->NW = c->P.n_samples;
    if(orig >= npx / 2 && (int)npx >= qvga) {
        const float sc = (float)npx / qvga;
        const int rawk = std::min((int)std::floor(0.5f + sc) + defk, defk + 4);
        c->median_k = (rawk % 2) ? rawk : rawk - 1;
        c->NG = maxG;
        for(auto& v : dsr) v |= 127;
    } else {
        const float sc = (float)orig / qvga;
        const int rawk = std::min((int)std::floor(0.5f + defk * sc * 2) + (defk - 4), defk);
        c->median_k = (rawk % 2) ? rawk : rawk - 1;
        c->NG = (int)std::min((size_t)std::pow((double)((float)maxG * sc), 2.0) + 1, (size_t)maxG);
    }
    if(c->median_k < 1) c->median_k = 1;
    if(C == 1) { c->NW = std::max(c->NW / 2, 1); c->NG = std::max(c->NG / 2, 1); }
    REQUIRE(c->NG <= PAW_MAXG, "too many global words");
    REQUIRE(c->NW <= PAW_SPLIT_MAX_NW, "PAWCS: at most 56 local words per pixel are supported (the reference's default is 50)");
    c->ds_roi_count = 0; for(uint8_t v : dsr) c->ds_roi_count += v != 0;
    REQUIRE(c->ds_roi_count > 0, "downsampled ROI is empty");
    f.median_k = c->median_k; f.auto_reset = 1;
    const int NW = c->NW, NG = c->NG;
    cudaStream_t st = c->stream;
    c->lw_key = dalloc<uint2>(st, (size_t)NW * c->plane, false);
    c->lw_rec = dalloc<uint8_t>(st, (size_t)NW * c->plane * c->paw_rec_bytes(), false);
    { // (first=1,last=0,occ=0): word not created yet
      std::vector<uint2> keys((size_t)NW * c->plane, make_uint2(0u, 1u)); h2d(st, c->lw_key, keys.data(), keys.size() * 8);
      std::vector<uint32_t> recs((size_t)NW * c->plane * (c->paw_rec_bytes() / 4), 0u);
      const size_t rw = c->paw_rec_bytes() / 4;
      for(size_t i = 0; i < (size_t)NW * c->plane; ++i) recs[i * rw + rw - 1] = 1u;
      h2d(st, c->lw_rec, recs.data(), recs.size() * 4); }
    c->gmap = dalloc<float>(st, (size_t)NG * c->gW * c->gH);
    c->gmap_tmp = dalloc<float>(st, (size_t)NG * c->gW * c->gH);
    c->gd = dalloc<GDict>(st, 1);
    { std::vector<uint8_t> l((size_t)NG * c->plane); for(int i = 0; i < NG; ++i) std::fill(l.begin() + (size_t)i * c->plane, l.begin() + (size_t)(i + 1) * c->plane, (uint8_t)i);
      c->glut = dalloc<uint8_t>(st, l.size(), false); h2d(st, c->glut, l.data(), l.size()); }
    c->paw_intents = dalloc<uint4>(st, c->plane, false);
    c->hand = dalloc<uint2>(st, c->plane);                 // split phase A: matched-word mask + flags per pixel
    c->wl2_idx = dalloc<uint32_t>(st, c->plane, false);    // work-list of the pixels whose scan goes past PAW_K words
    c->wl_ctx = dalloc<uint32_t>(st, c->plane * 2, false); // phase B work-list: (target pixel, remaining hits)
    c->gop_w = dalloc<float>(st, c->plane, false);
    c->gop_g = dalloc<uint8_t>(st, c->plane, false);
    c->ds_roi = dalloc<uint8_t>(st, dsr.size(), false); h2d(st, c->ds_roi, dsr.data(), dsr.size());
    c->bgimg = dalloc<uint8_t>(st, npx * C);
    c->maps = dalloc<float4>(st, c->plane * 2);
    c->fin = dalloc<float2>(st, c->plane);
    c->dsLT = dalloc<float>(st, (size_t)c->dsW * c->dsH * C);
    c->dsST = dalloc<float>(st, (size_t)c->dsW * c->dsH * C);
    { std::vector<float4> m(c->plane * 2);
      for(size_t i = 0; i < c->plane; ++i) { m[i * 2] = make_float4(1.0f, 2.0f, 10.0f, 0.f); m[i * 2 + 1] = make_float4(0.f, 0.f, 0.f, 0.f); } // T, R, v (:473-477)
      h2d(st, c->maps, m.data(), m.size() * sizeof(float4)); }
    { std::vector<uint32_t> rb(bp, 0);
      for(int y = 0; y < H; ++y) for(int x = 0; x < W; ++x) if(c->roi_host[(size_t)y * W + x] == 255) rb[(size_t)y * c->WW + (x >> 5)] |= 1u << (x & 31);
      h2d(st, c->roi255, rb.data(), bp * 4); }
    GDict g{};
    for(int i = 0; i < PAW_MAXG; ++i) g.dict[i] = -1;
    g.weight_offset = PAW_WEIGHT_OFFSET; g.boot = 1; g.rep_winner = 0xFFFFFFFFu; g.g_rep = -1;
    g.ds_roi_count = (uint32_t)c->ds_roi_count; g.nST = ((uint32_t)c->P.n_samples_for_moving_avgs / 2u) / 4u;
    g.refresh_req = PAW_REQ_REFRESH; g.refresh_base_occ = 1; g.refresh_decr = 0.0f; g.refresh_force = 0; // refreshModel(1,0) (:555)
    h2d(st, c->gd, &g, sizeof(g));
    c->paw_frame = 1;
}

PawArgs paw_args(lvb_context* c, const uint8_t* img, size_t pitch, int use_tma, double lr) {
    PawArgs A{};
    A.W = c->W; A.H = c->H; A.Wp = c->Wp; A.WW = c->WW; A.NW = c->NW; A.NG = c->NG; A.gW = c->gW; A.gH = c->gH; A.plane = c->plane;
    A.img = img; A.ipitch = pitch;
    A.lw_key = c->lw_key; A.lw_rec = c->lw_rec;
    A.glut = c->glut; A.gmap = c->gmap; A.gmap_tmp = c->gmap_tmp; A.gd = c->gd;
    A.maps = c->maps; A.fin = c->fin; A.last_color = c->last_color; A.last_desc = c->last_desc;
    A.roi_bits = c->roi_bits; A.roi255_bits = c->roi255; A.raw_bits = c->raw; A.unstable_bits = c->unstable; A.blinks_bits = c->blinks; A.lastfg_bits = c->lastfg;
    A.illum_bits = c->illum; A.did_bits = c->did; A.dil_bits = c->dil; A.dilinv_bits = c->dilinv;
    A.intent_bits = c->intent_bits; A.intents = c->paw_intents; A.bitplane = (size_t)c->H * c->WW;
    A.gop_bits = c->gop_bits; A.gop_w = c->gop_w; A.gop_g = c->gop_g;
    A.hand = c->hand; A.wl = c->wl2_idx; A.wlB = (uint2*)c->wl_ctx;
    A.lut = c->lut; A.ctl = c->ctl; A.seed = c->seed; A.lr_fixed = lr_to_fixed(lr);
    A.min_color = c->P.color_dist_threshold; A.desc_off = c->P.desc_dist_threshold; A.use_tma = use_tma; A.collect_stats = c->collect_stats;
    A.rel = c->P.rel_lbsp_threshold; A.lbsp_off = c->P.lbsp_threshold_offset; A.avg_samples = c->P.n_samples_for_moving_avgs;
    A.dsW = c->dsW; A.dsH = c->dsH; A.ds_roi = c->ds_roi; A.dsLT = c->dsLT; A.dsST = c->dsST; A.bgimg = c->bgimg;
    return A;
}

/// host-requested BackgroundSubtractorPAWCS::refreshModel(nBaseOccCount, fOccDecrFrac, bForceFGUpdate)
void paw_request_refresh(lvb_context* c, uint32_t base_occ, float decr, bool force) {
    REQUIRE(c->initialized, "algo must be initialized first");
    REQUIRE(decr >= 0.0f && decr <= 1.0f, "model occurrence decrementation must be given as a non-null fraction");
    GDict g;
    CK(cudaMemcpyAsync(&g, c->gd, sizeof(g), cudaMemcpyDeviceToHost, c->stream)); CK(cudaStreamSynchronize(c->stream));
    g.refresh_req = PAW_REQ_REFRESH; g.refresh_base_occ = base_occ; g.refresh_decr = decr; g.refresh_force = force ? 1 : 0; g.set_T_one = 0;
    CK(cudaMemcpyAsync(c->gd, &g, sizeof(g), cudaMemcpyHostToDevice, c->stream)); CK(cudaStreamSynchronize(c->stream));
    paw_launch_refresh(c, paw_args(c, c->d_img, c->ipitch, c->use_tma, 0.0), 1);
    CK(cudaStreamSynchronize(c->stream));
}

/// host-requested refreshModel(frac, force): fill the request in FrameCtl, then the same kernels the frame tail uses
void request_refresh(lvb_context* c, float frac, bool force) {
    REQUIRE(c->initialized, "algo must be initialized first");
    REQUIRE(frac > 0.0f && frac <= 1.0f, "model refresh must be given as a non-null fraction");
    FrameCtl f; get_ctl(c, f);
    const uint32_t N = (uint32_t)c->P.n_samples;
    f.do_refresh = 1; f.set_T_one = 0; f.refresh_force = force ? 1 : 0;
    f.refresh_count = frac < 1.0f ? (uint32_t)(frac * (float)N) : N;
    f.refresh_start = 0;
    put_ctl(c, f);
    if(frac < 1.0f) { refresh_start_kernel<<<1, 1, 0, c->stream>>>(c->ctl, c->seed, N); LAUNCHED(); }
    launch_refresh(c);
    CK(cudaStreamSynchronize(c->stream));
}

void sync(lvb_context* c);
void do_initialize(lvb_context* c, const uint8_t* img, int W, int H, int C, size_t step, const uint8_t* roi) {
    REQUIRE(img && W > 0 && H > 0 && (C == 1 || C == 3 || C == 4), "provided image for initialization must be non-empty, continuous, and of type 8UC1/3/4");
    REQUIRE(C != 4, "8UC4 input is accepted by the reference's initialize() but its apply() has no 4-channel path; use 8UC1 or 8UC3");
    REQUIRE(W >= 5 && H >= 5, "image too small for the 5x5 LBSP pattern");
    REQUIRE(W <= 8192, "frame width above 8192 pixels is not supported");
    REQUIRE(step >= (size_t)W * C, "row step smaller than a row");
    CK(cudaSetDevice(c->device));
    if(c->initialized) sync(c);   // frames still in flight (lvb_apply_async) are collected, their masks delivered, before the buffers go
    // ROI (BackgroundSubtractionUtils.cpp:82-99, validateROI :28-36)
    std::vector<uint8_t> r((size_t)W * H, 255);
    if(roi) {
        for(size_t i = 0; i < (size_t)W * H; ++i) REQUIRE(roi[i] == 0 || roi[i] == 255, "provided ROI mat values must be 0 or 255 only");
        std::vector<uint8_t> dil((size_t)W * H);
        morph_rect_host(roi, dil.data(), W, H, 2);
        for(size_t i = 0; i < (size_t)W * H; ++i) r[i] = roi[i] | (dil[i] ? 128 : 0); // 255/2 saturate-rounds to 128
    } else if(c->roi_host.size() == (size_t)W * H && c->W == W && c->H == H) r = c->roi_host; // reuse last ROI if sizes match (:87-88)
    size_t orig = 0, fin = 0;
    for(uint8_t v : r) orig += v != 0;
    REQUIRE(orig > 0, "provided ROI mat contains no useful pixels");
    for(int y = 0; y < H; ++y) for(int x = 0; x < W; ++x) if(x < 2 || y < 2 || x >= W - 2 || y >= H - 2) r[(size_t)y * W + x] = 0;
    for(uint8_t v : r) fin += v != 0;
    REQUIRE(fin > 0, "provided ROI mat contains no useful pixels away from borders (descriptors will hit image bounds)");

    c->free_all();
    c->W = W; c->H = H; c->C = C; c->Wp = (W + 31) / 32 * 32; c->WW = c->Wp / 32; c->plane = (size_t)H * c->Wp;
    c->dsW = W / 8; c->dsH = H / 8;
    c->roi_host = r; c->orig_roi_count = orig; c->roi_count = fin;
    const int N = c->P.n_samples;
    c->ipitch = ((size_t)W * C + 127) / 128 * 128;
    c->d_img = dalloc<uint8_t>(c->stream, c->ipitch * H);
    c->d_mask = dalloc<uint8_t>(c->stream, (size_t)W * H);
    CK(cudaMallocHost((void**)&c->h_img, (size_t)W * H * C));
    CK(cudaMallocHost((void**)&c->h_mask, (size_t)W * H));
    c->use_tma = make_image_tmap(&c->tmap_img, c->d_img, W, H, C, c->ipitch) ? 1 : 0;
    c->ext_ptr = nullptr;
    {   // pipeline slots: slot 0 aliases the buffers above, slot 1 gets its own
        lvb_context::Slot& a = c->slot[0]; lvb_context::Slot& b = c->slot[1];
        a.d_img = c->d_img; a.d_mask = c->d_mask; a.h_img = c->h_img; a.h_mask = c->h_mask; a.tmap = c->tmap_img; a.use_tma = c->use_tma;
        a.used = b.used = false;
        b.d_img = dalloc<uint8_t>(c->stream, c->ipitch * H); b.d_mask = dalloc<uint8_t>(c->stream, (size_t)W * H);
        CK(cudaMallocHost((void**)&b.h_img, (size_t)W * H * C)); CK(cudaMallocHost((void**)&b.h_mask, (size_t)W * H));
        b.use_tma = make_image_tmap(&b.tmap, b.d_img, W, H, C, c->ipitch) ? 1 : 0;
        if(!c->s_in) { CK(cudaStreamCreateWithFlags(&c->s_in, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&c->s_out, cudaStreamNonBlocking));
                       { int lo = 0, hi = 0; CK(cudaDeviceGetStreamPriorityRange(&lo, &hi)); CK(cudaStreamCreateWithPriority(&c->s_aux, cudaStreamNonBlocking, lo)); }
                       { int lo = 0, hi = 0; CK(cudaDeviceGetStreamPriorityRange(&lo, &hi)); CK(cudaStreamCreateWithPriority(&c->s_post, cudaStreamNonBlocking, hi)); }
                       for(cudaEvent_t* e : {&c->ev_scan, &c->ev_post, &c->ev_ds, &c->ev_mask}) CK(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
                       CK(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming)); }
        for(lvb_context::Slot& sl : c->slot) if(!sl.h2d_done) {
            CK(cudaEventCreateWithFlags(&sl.h2d_done, cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&sl.compute_done, cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&sl.d2h_done, cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&sl.in_consumed, cudaEventDisableTiming));
        }
    }
    if(c->algo != LVB_ALGO_PAWCS) {
        c->bg = dalloc<uint8_t>(c->stream, (size_t)N * c->plane * c->rec_bytes());
    }
    c->last_color = dalloc<uint8_t>(c->stream, c->plane * c->col_bytes());
    c->last_desc = dalloc<uint8_t>(c->stream, c->plane * c->desc_bytes());
    c->tmp_desc = dalloc<uint8_t>(c->stream, c->plane * c->desc_bytes());
    const size_t bp = (size_t)H * c->WW;
    c->bits = dalloc<uint32_t>(c->stream, bp * 27);
    c->raw_alt = c->bits + bp * 24; c->blinks_alt = c->bits + bp * 25; c->lastfg_alt = c->bits + bp * 26;
    uint32_t** planes[] = {&c->roi_bits, &c->raw, &c->lastraw, &c->lastrawblink, &c->blinks, &c->tmpA, &c->pre, &c->reach, &c->comb,
                           &c->lastfg, &c->dilinv, &c->unstable, &c->ghost[0], &c->ghost[1], &c->intent_bits};
    for(int i = 0; i < 15; ++i) *planes[i] = c->bits + bp * i;
    c->roi255 = c->bits + bp * 19; c->illum = c->bits + bp * 20; c->did = c->bits + bp * 21; c->dil = c->bits + bp * 22; c->gop_bits = c->bits + bp * 23; // PAWCS (intent planes: 14..18)
    c->ghost_idx = 0; c->sub_frame = 1; c->chain_seq = 0; c->post_pending = false; c->fin_pending = 0;
    c->uf_rs = (W + 1) / 2 + 1;
    c->uf_parent = dalloc<uint32_t>(c->stream, (size_t)H * c->uf_rs + 1);
    c->uf_rankbase = dalloc<ushort>(c->stream, (size_t)H * c->WW);
    c->intents = dalloc<ushort>(c->stream, c->plane);
    c->lut = dalloc<uint8_t>(c->stream, 256);
    c->ctl = dalloc<FrameCtl>(c->stream, 1);
    {   // bit-packed ROI
        std::vector<uint32_t> rb(bp, 0);
        for(int y = 0; y < H; ++y) for(int x = 0; x < W; ++x) if(r[(size_t)y * W + x]) rb[(size_t)y * c->WW + (x >> 5)] |= 1u << (x & 31);
        h2d(c->stream, c->roi_bits, rb.data(), bp * 4);
    }
    {   // LBSP threshold LUT (BackgroundSubtractorLBSP.cpp:29-30, 42-43; quirk Q2)
        uint8_t lut[256];
        for(int t = 0; t < 256; ++t) {
            const float v = C == 1 ? ((float)t * c->P.rel_lbsp_threshold + (float)c->P.lbsp_threshold_offset) / 3 : (float)t * c->P.rel_lbsp_threshold + (float)c->P.lbsp_threshold_offset;
            const long q = std::lrint((double)v);
            lut[t] = (uint8_t)(q < 0 ? 0 : q > 255 ? 255 : q);
        }
        h2d(c->stream, c->lut, lut, 256);
        c->lut_small = lut_fits_7bit(c, lut);
    }
    {   // floor(2^32 / n) for n = 1..256 (n = 1: 2^32 - 1, still exact with fast_mod's single correction step)
        uint32_t mg[257];
        mg[0] = 0; mg[1] = 0xFFFFFFFFu;
        for(uint32_t n = 2; n <= 256; ++n) mg[n] = (uint32_t)(0x100000000ull / n);
        c->magic = dalloc<uint32_t>(c->stream, 257, false);
        h2d(c->stream, c->magic, mg, sizeof(mg));
        // i / colorRange, i / descRange (SuBSENSE.cpp:57-60: 255 / 16 for 1 channel, 765 / 48 for 3): float division on the host
        // is the same IEEE round-to-nearest quotient as __fdiv_rn
        const int cr = C == 1 ? 255 : 765, dr = C == 1 ? 16 : 48;
        std::vector<float> dv((size_t)cr + 1 + dr + 1);
        for(int i = 0; i <= cr; ++i) dv[i] = (float)i / (float)cr;
        for(int i = 0; i <= dr; ++i) dv[(size_t)cr + 1 + i] = (float)i / (float)dr;
        c->div_tab = dalloc<float>(c->stream, dv.size(), false);
        h2d(c->stream, c->div_tab, dv.data(), dv.size() * sizeof(float));
    }
    FrameCtl f{};
    f.frame_idx = 1; f.aLT = 1.0f; f.aST = 1.0f; f.roi_count = (uint32_t)fin;
    f.auto_reset = 1; f.median_k = c->P.median_blur_kernel_size;
    if(c->algo == LVB_ALGO_SUBSENSE) { // SuBSENSE.cpp:112-128
        const int tot = W * H, qvga = 320 * 240;
        if(orig >= (size_t)tot / 2 && tot >= qvga) {
            f.lr_scaling = 1; f.auto_reset = 1; f.use3x3 = !(tot > qvga * 2);
            const int rawk = std::min((int)std::floor((float)tot / qvga + 0.5f) + c->P.median_blur_kernel_size, 14);
            f.median_k = (rawk % 2) ? rawk : rawk - 1;
            f.t_lower = 2.0f; f.t_upper = 256.0f;
        } else {
            f.lr_scaling = 0; f.auto_reset = 0; f.use3x3 = 1; f.median_k = c->P.median_blur_kernel_size;
            f.t_lower = 4.0f; f.t_upper = 512.0f;
        }
        c->maps = dalloc<float4>(c->stream, c->plane * 2);
        c->fin = dalloc<float2>(c->stream, c->plane);
        c->hand = dalloc<uint2>(c->stream, c->plane);
        c->own_slot = dalloc<uint8_t>(c->stream, c->plane, false);
        CK(cudaMemsetAsync(c->own_slot, 0xFF, c->plane, c->stream));
        if(!getenv("LVB_NO_CBOX")) {   // starts as "contains everything" (never filters); initialize()'s refreshModel makes it exact
            c->cbox = dalloc<uint2>(c->stream, c->plane, false);
            std::vector<uint2> full(c->plane, make_uint2(0u, 0x00FFFFFFu)); h2d(c->stream, c->cbox, full.data(), full.size() * sizeof(uint2));
        }
        c->wl_cap = (uint32_t)c->plane;   // every pixel may be undecided after two samples (first frames after a scene change)
        c->wl_ctx = dalloc<uint32_t>(c->stream, (size_t)(C == 1 ? WlCtx<1>::FIELDS : WlCtx<3>::FIELDS) * c->wl_cap, false);
        c->wl2_idx = dalloc<uint32_t>(c->stream, c->wl_cap, false);
        c->last_color_alt = dalloc<uint8_t>(c->stream, c->plane * c->col_bytes());
        c->last_desc_alt = dalloc<uint8_t>(c->stream, c->plane * c->desc_bytes());
        c->dsLT = dalloc<float>(c->stream, (size_t)c->dsW * c->dsH * C);
        c->dsST = dalloc<float>(c->stream, (size_t)c->dsW * c->dsH * C);
        std::vector<float4> m(c->plane * 2);
        for(size_t i = 0; i < c->plane; ++i) { m[i * 2] = make_float4(f.t_lower, 1.0f, 10.0f, 0.f); m[i * 2 + 1] = make_float4(0.f, 0.f, 0.f, 0.f); }
        h2d(c->stream, c->maps, m.data(), m.size() * sizeof(float4));
        std::vector<float> r1(c->plane, 1.0f);
        c->r_plane = dalloc<float>(c->stream, c->plane, false);
        h2d(c->stream, c->r_plane, r1.data(), r1.size() * sizeof(float));
        {   // TMA maps of the planes the scan kernel stages (subsense.cuh: ScanTile). Pair i describes colour / descriptor planes i as "previous frame".
            const int rec = (int)c->rec_bytes(), bge = rec >= 8 ? 8 : 4, IH = TILE_H + 2 * HALO;
            const int bw_col = C == 1 ? ScanTile<1>::BW_COL : ScanTile<3>::BW_COL, bw_desc = C == 1 ? ScanTile<1>::BW_DESC : ScanTile<3>::BW_DESC;
            const int bw_int = C == 1 ? ScanTile<1>::BW_INT : ScanTile<3>::BW_INT;
            void* cols[2] = {c->last_color, c->last_color_alt}; void* descs[2] = {c->last_desc, c->last_desc_alt};
            bool ok = true;
            for(int i = 0; i < 2; ++i) {
                ScanMaps& M = c->scan_maps[i];
                ok = ok && make_plane_tmap(&M.pcol, cols[i], (int)c->col_bytes(), c->Wp, H, bw_col, IH);
                ok = ok && make_plane_tmap(&M.pdesc, descs[i], (int)c->desc_bytes(), c->Wp, H, bw_desc, IH);
                ok = ok && make_plane_tmap(&M.intents, c->intents, 2, c->Wp, H, bw_int, IH);
                ok = ok && make_plane_tmap(&M.own, c->own_slot, 1, c->Wp, H, TILE_W, TILE_H);
                ok = ok && make_plane_tmap(&M.rpl, c->r_plane, 4, c->Wp, H, TILE_W, TILE_H);
                ok = ok && make_plane_tmap(&M.bg, c->bg, bge, (size_t)c->Wp * (rec / bge), (size_t)N * H, TILE_W * (rec / bge), TILE_H);
            }
            REQUIRE(ok, "cuTensorMapEncodeTiled failed for the SuBSENSE scan planes (driver without TMA support?)");
            c->scan_maps_idx = 0;
        }
    }
    if(c->algo == LVB_ALGO_PAWCS) paw_initialize(c, f, orig);
    c->median_k = f.median_k;
    REQUIRE(c->median_k >= 1 && c->median_k <= 31 && (c->median_k & 1), "median blur kernel size must be odd and <= 31");
    // first refresh request: all N slots from slot 0 (SuBSENSE.cpp:184 refreshModel(1.0f); LOBSTER.cpp:455 refreshModel(1.0f,true))
    f.do_refresh = 1; f.refresh_epoch = 0; f.refresh_start = 0; f.refresh_count = (uint32_t)N; f.refresh_force = c->algo == LVB_ALGO_LOBSTER ? 1 : 0;
    h2d(c->stream, c->ctl, &f, sizeof(f));
    // frame upload + init kernel
    CK(cudaMemcpy2DAsync(c->d_img, c->ipitch, img, step, (size_t)W * C, H, cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    InitArgs I{};
    I.W = W; I.H = H; I.Wp = c->Wp; I.WW = c->WW; I.img = c->d_img; I.ipitch = c->ipitch; I.last_color = c->last_color; I.last_desc = c->last_desc;
    I.roi_bits = c->roi_bits; I.lut = c->lut; I.use_tma = c->use_tma;
    if(C == 1) init_frame_kernel<1><<<stage_grid(c), stage_block, 0, c->stream>>>(I, c->tmap_img);
    else init_frame_kernel<3><<<stage_grid(c), stage_block, 0, c->stream>>>(I, c->tmap_img);
    LAUNCHED();
    if(c->last_color_alt) { // pixels outside the ROI keep their initial colour / descriptor in both planes
        CK(cudaMemcpyAsync(c->last_color_alt, c->last_color, c->plane * c->col_bytes(), cudaMemcpyDeviceToDevice, c->stream));
        CK(cudaMemcpyAsync(c->last_desc_alt, c->last_desc, c->plane * c->desc_bytes(), cudaMemcpyDeviceToDevice, c->stream));
    }
    c->initialized = true;
    if(c->algo == LVB_ALGO_PAWCS) paw_launch_refresh(c, paw_args(c, c->d_img, c->ipitch, c->use_tma, 0.0), 1);
    else launch_refresh(c);
    CK(cudaStreamSynchronize(c->stream));
    c->stat_frames = 0;
}

uint32_t lr_to_fixed(double lr) {
    if(std::isinf(lr)) return 0xFFFFFFFFu;
    if(lr > 0) { const double v = std::ceil(lr); return v >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)v; }
    return 0;
}

/// one PAWCS frame on the instance's stream (kernel order: see pawcs.cuh)
void paw_enqueue_frame(lvb_context* c, const uint8_t* img, size_t pitch, const CUtensorMap& tmap, int use_tma, uint8_t* d_mask_out, double lr) {
    const int W = c->W, H = c->H, C = c->C;
    cudaStream_t st = c->stream;
    const PawArgs A = paw_args(c, img, pitch, use_tma, lr);
    const dim3 tg = tile_grid(c), tb(32, 8), wg = word_grid(c), mg(c->Wp / 32, (H + 8 * MEDIAN_ROWS - 1) / (8 * MEDIAN_ROWS));
    const uint32_t frame = c->paw_frame;
    const bool boot = frame <= PAW_BOOTSTRAP;
    const uint32_t grate = boot ? 8u : 16u;
    const int recalc = (frame % (grate << 5)) == 0, update = (frame % grate) == 0, check_model = (frame % PAW_BOOTSTRAP) == 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    if(c->profile) { CK(cudaEventCreate(&ev0)); CK(cudaEventCreate(&ev1)); CK(cudaEventRecord(ev0, st)); }
    // instance stream : scan -> scan tail (work-list pixels, incl. their bubble pass) -> global dictionary -> mask chain -> ...
    // auxiliary stream: bubble pass of the other pixels (beside the scan tail) -> phase B (needs both) -> phase B tail
    {
        const int tail_grid = c->sm_count * PAWT_MIN_BLOCKS;
        if(c->lut_small) { if(C == 1) pawcs_scan<1, true><<<stage_grid(c), stage_block, 0, st>>>(A, tmap); else pawcs_scan<3, true><<<stage_grid(c), stage_block, 0, st>>>(A, tmap); }
        else { if(C == 1) pawcs_scan<1, false><<<stage_grid(c), stage_block, 0, st>>>(A, tmap); else pawcs_scan<3, false><<<stage_grid(c), stage_block, 0, st>>>(A, tmap); }
        LAUNCHED();
        CK(cudaEventRecord(c->ev_fork, st)); CK(cudaStreamWaitEvent(c->s_aux, c->ev_fork, 0));
        if(C == 1) pawcs_bubble<1><<<tg, tb, 0, c->s_aux>>>(A); else pawcs_bubble<3><<<tg, tb, 0, c->s_aux>>>(A);
        LAUNCHED();
        if(c->lut_small) { if(C == 1) pawcs_scan_tail<1, true><<<tail_grid, PAW_TAIL_THREADS, 0, st>>>(A); else pawcs_scan_tail<3, true><<<tail_grid, PAW_TAIL_THREADS, 0, st>>>(A); }
        else { if(C == 1) pawcs_scan_tail<1, false><<<tail_grid, PAW_TAIL_THREADS, 0, st>>>(A); else pawcs_scan_tail<3, false><<<tail_grid, PAW_TAIL_THREADS, 0, st>>>(A); }
        LAUNCHED();
        CK(cudaEventRecord(c->ev_scan, st)); CK(cudaStreamWaitEvent(c->s_aux, c->ev_scan, 0));
    }
    if(c->profile) { CK(cudaEventRecord(ev1, st)); c->prof_events.push_back(ev0); c->prof_events.push_back(ev1); }
    if(C == 1) pawcs_phaseB<1><<<tg, tb, 0, c->s_aux>>>(A); else pawcs_phaseB<3><<<tg, tb, 0, c->s_aux>>>(A);
    LAUNCHED();
    if(C == 1) pawcs_phaseB_tail<1><<<c->sm_count * PAWBT_MIN_BLOCKS, PAW_TAIL_THREADS, 0, c->s_aux>>>(A); else pawcs_phaseB_tail<3><<<c->sm_count * PAWBT_MIN_BLOCKS, PAW_TAIL_THREADS, 0, c->s_aux>>>(A);
    LAUNCHED();
    CK(cudaEventRecord(c->ev_join, c->s_aux));
    if(C == 1) pawcs_gword_replace<1><<<16, 1024, 0, st>>>(A); else pawcs_gword_replace<3><<<16, 1024, 0, st>>>(A);
    LAUNCHED();
    pawcs_gword_apply<<<dim3((c->gW + 31) / 32, (c->gH + 7) / 8), 256, 0, st>>>(A); LAUNCHED();
    pawcs_gword_finish<<<1, 128, 0, st>>>(A, !(recalc || update)); LAUNCHED();
    if(recalc || update) {
        const dim3 gmg((c->gW * c->gH + PAW_GM_THREADS * PAW_GM_PER_THREAD - 1) / (PAW_GM_THREADS * PAW_GM_PER_THREAD), c->NG);
        if(recalc) { pawcs_gmaint_sum<<<gmg, PAW_GM_THREADS, 0, st>>>(A); LAUNCHED(); }
        pawcs_gmaint_weights<<<1, PAW_MAXG, 0, st>>>(A, recalc, update); LAUNCHED();
        pawcs_gmaint_blur<<<gmg, PAW_GM_THREADS, 0, st>>>(A, update); LAUNCHED();
        if(update) { pawcs_gmaint_copy<<<gmg, PAW_GM_THREADS, 0, st>>>(A); LAUNCHED(); }
        pawcs_gdict_bubble<<<1, 1, 0, st>>>(A); LAUNCHED();
    }
    if(update) { pawcs_glut_bubble<<<tg, tb, 0, st>>>(A, 0); LAUNCHED(); }
    PostArgs P{};
    P.W = W; P.H = H; P.WW = c->WW; P.Wp = c->Wp; P.raw = c->raw; P.lastraw = c->lastraw; P.lastrawblink = c->lastrawblink; P.blinks = c->blinks;
    P.tmpA = c->tmpA; P.pre = c->pre; P.reach = c->reach; P.comb = c->comb; P.lastfg = c->lastfg; P.dilinv = c->dilinv; P.dil = c->dil;
    P.out_mask = d_mask_out; P.out_pitch = (size_t)W; P.fin = c->fin; P.fin_in = c->fin; P.ctl = c->ctl; P.median_k = c->median_k;
    P.did = c->did; P.roi255 = c->roi255; P.roi = c->roi_bits; P.illum = c->illum; // + next frame's illumination mask
    pp_blink_close<<<wg, 256, 0, st>>>(P); LAUNCHED();
    {
        HoleArgs Hh{};
        Hh.W = W; Hh.H = H; Hh.WW = c->WW; Hh.RS = c->uf_rs; Hh.pre = c->pre; Hh.raw = c->raw; Hh.comb = c->comb;
        Hh.parent = c->uf_parent; Hh.rankbase = c->uf_rankbase;
        const int rb = (H + 7) / 8;
        pp_holes_init<<<rb, 256, 0, st>>>(Hh); LAUNCHED();
        pp_holes_union<<<rb, 256, 0, st>>>(Hh); LAUNCHED();
        pp_holes_combine<<<rb, 256, 0, st>>>(Hh); LAUNCHED();
    }
    pp_median<<<mg, tb, 0, st>>>(c->comb, c->lastfg, d_mask_out, (size_t)W, W, H, c->WW, c->median_k); LAUNCHED();
    pp_dilate_blink<<<wg, 256, 0, st>>>(P); LAUNCHED();
    pp_final_ema<<<tg, tb, 0, st>>>(P); LAUNCHED();
    const int nds = c->dsW * c->dsH;
    if(C == 1) pawcs_motion_kernel<1><<<(nds + 127) / 128, 128, 0, st>>>(A); else pawcs_motion_kernel<3><<<(nds + 127) / 128, 128, 0, st>>>(A);
    LAUNCHED();
    CK(cudaStreamWaitEvent(st, c->ev_join, 0)); // everything below may read or rewrite the local dictionaries
    if(check_model) {
        if(C == 1) pawcs_background_kernel<1><<<tg, tb, 0, st>>>(A, c->bgimg, nullptr, 0); else pawcs_background_kernel<3><<<tg, tb, 0, st>>>(A, c->bgimg, nullptr, 0);
        LAUNCHED();
        if(C == 1) pawcs_model_dist_kernel<1><<<(nds + 127) / 128, 128, 0, st>>>(A); else pawcs_model_dist_kernel<3><<<(nds + 127) / 128, 128, 0, st>>>(A);
        LAUNCHED();
    }
    pawcs_tail1_kernel<<<1, 256, 0, st>>>(A, check_model, !check_model); LAUNCHED();
    if(check_model) {
        paw_launch_refresh(c, A, 0);                  // moving-camera mode switch (:1486-1499)
        pawcs_tail2_kernel<<<1, 1, 0, st>>>(A); LAUNCHED();
    }
    paw_launch_refresh(c, A, 1);                      // frame-level model reset (:1503-1510)
    c->paw_frame = frame + 1;
    if(c->collect_stats) ++c->stat_frames;
}

/// enqueue one frame on the instance's stream; the frame is already in device memory at (img,pitch)
/// `mask_ready` (optional) is recorded on whichever stream produces d_mask_out, right behind the kernel that writes it
void enqueue_frame(lvb_context* c, const uint8_t* img, size_t pitch, const CUtensorMap& tmap, int use_tma, uint8_t* d_mask_out, double lr, cudaEvent_t mask_ready = nullptr) {
    if(c->algo == LVB_ALGO_PAWCS) { paw_enqueue_frame(c, img, pitch, tmap, use_tma, d_mask_out, lr); if(mask_ready) CK(cudaEventRecord(mask_ready, c->stream)); return; }
    const int W = c->W, H = c->H, C = c->C;
    cudaStream_t st = c->stream;
    const bool sub = c->algo == LVB_ALGO_SUBSENSE;
    SubArgs A{};
    A.W = W; A.H = H; A.Wp = c->Wp; A.WW = c->WW; A.N = c->P.n_samples; A.REQ = c->P.n_required; A.plane = c->plane;
    A.img = img; A.ipitch = pitch; A.bg = c->bg; A.maps = c->maps; A.fin = c->fin; A.hand = c->hand;
    A.ema_frame = sub ? c->fin_pending : 0u; A.avg_samples = c->P.n_samples_for_moving_avgs;
    A.last_color = sub ? c->last_color_alt : c->last_color; A.last_desc = sub ? c->last_desc_alt : c->tmp_desc;
    A.prev_color = c->last_color; A.prev_desc = c->last_desc; A.pending_seq = sub ? c->nb_seq : 0u; A.roi_bits = c->roi_bits; A.raw_bits = sub ? c->raw_alt : c->raw; A.unstable_bits = c->unstable;
    A.blinks_bits = c->blinks; A.lastfg_bits = c->lastfg; A.ghost_prev = c->ghost[c->ghost_idx]; A.ghost_cur = c->ghost[c->ghost_idx ^ 1];
    A.intents = c->intents; A.lut = c->lut; A.ctl = c->ctl; A.seed = c->seed;
    A.wl_ctx = c->wl_ctx; A.wl_cap = c->wl_cap; A.wl2_idx = c->wl2_idx; A.own_slot = c->own_slot; A.cbox = sub ? c->cbox : nullptr;
    A.lr_fixed = lr_to_fixed(lr); A.min_color = c->P.color_dist_threshold; A.desc_off = c->P.desc_dist_threshold;
    A.use_tma = use_tma; A.collect_stats = c->collect_stats;
    A.n_magic = (uint32_t)(0x100000000ull / (uint64_t)c->P.n_samples);
    A.r_plane = c->r_plane; A.div_color = c->div_tab; A.div_desc = c->div_tab + (C == 1 ? 256 : 766);
    A.magic = c->magic; A.lr_magic = magic_of(A.lr_fixed); A.lr2_magic = magic_of(A.lr_fixed / 2u + 1u);
    PhaseBArgs B{};
    B.W = W; B.H = H; B.Wp = c->Wp; B.WW = c->WW; B.CH = C; B.plane = c->plane;
    B.bg = c->bg; B.last_color = c->last_color; B.last_desc = A.last_desc; B.intents = c->intents; // LOBSTER
    const dim3 tg = tile_grid(c), tb(32, 8), wg = word_grid(c), mg(c->Wp / 32, (H + 8 * MEDIAN_ROWS - 1) / (8 * MEDIAN_ROWS));

    auto mark = [&](cudaStream_t on, const char* n) { if(c->trace_on && c->profile) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, on); c->trace.push_back({n, e}); } };
    mark(st, "frame_start");
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    if(c->profile) { CK(cudaEventCreate(&ev0)); CK(cudaEventCreate(&ev1)); CK(cudaEventRecord(ev0, st)); }
    if(sub) {
        // Three streams per frame k (see subsense.cuh):
        //   instance stream : scan(k) -> [chain(k-1) done, motion(k) done] -> feedback(k) + frame tail -> phase B(k) -> conditional refresh(k)
        //   mask stream     : [scan(k) done] -> blink/close -> holes -> median (mask out) -> completion counter -> dilate/blink -> final EMAs
        //   auxiliary stream: frame-level motion analysis (needs only the input frame)
        cudaStream_t sp = c->s_post;
        const uint32_t seq = ++c->chain_seq;
        CK(cudaEventRecord(c->ev_fork, st)); CK(cudaStreamWaitEvent(c->s_aux, c->ev_fork, 0));
        {
            DownsampleArgs D{};
            D.W = W; D.H = H; D.CH = C; D.dsW = c->dsW; D.dsH = c->dsH; D.img = img; D.ipitch = pitch; D.dsLT = c->dsLT; D.dsST = c->dsST; D.ctl = c->ctl;
            const int nds = c->dsW * c->dsH;
            if(nds > 0) {
                if(C == 1) downsample_motion_kernel<1><<<(nds + 127) / 128, 128, 0, c->s_aux>>>(D); else downsample_motion_kernel<3><<<(nds + 127) / 128, 128, 0, c->s_aux>>>(D);
                LAUNCHED();
            }
        }
        CK(cudaEventRecord(c->ev_ds, c->s_aux));
        const ScanMaps& SM = c->scan_maps[c->scan_maps_idx];   // "previous frame" = the latest colour / descriptor planes
        // multi-tile CTAs: every CTA walks ~3-4 tiles with the next tile's TMA boxes in flight. The grid is 4x what fits on the chip at once
        // (measured, 1080p: 4 / 8 / 16 / 32 / 55 CTAs per SM -> 134 / 117 / 113 / 117 / 129 us; one tile per CTA: 127 us): a fully persistent
        // grid loses to the tail effect of its static tile assignment, the oversubscribed one lets the hardware scheduler balance it
        static const int scan_ctas = getenv("LVB_SCAN_CTAS") ? atoi(getenv("LVB_SCAN_CTAS")) : SCAN_GRID_CTAS_PER_SM;
        const dim3 sg((unsigned)std::min<int>((int)(stage_grid(c).x * stage_grid(c).y), c->sm_count * scan_ctas));
        // the colour boxes only grow between rebuilds (every sample write widens them): an exact rebuild every 128 frames (one pass over the
        // model, ~0.3 ms at 1080p = 2 us per frame amortised) keeps them tight. Queued writes not yet in the model grow them when the scan applies them.
        if(c->cbox && (c->sub_frame % 128u) == 0u) rebuild_cbox(c);
        if(c->lut_small) { if(C == 1) subsense_scan<1, true><<<sg, stage_block, 0, st>>>(A, tmap, SM); else subsense_scan<3, true><<<sg, stage_block, 0, st>>>(A, tmap, SM); }
        else { if(C == 1) subsense_scan<1, false><<<sg, stage_block, 0, st>>>(A, tmap, SM); else subsense_scan<3, false><<<sg, stage_block, 0, st>>>(A, tmap, SM); }
        LAUNCHED();
        if(c->profile) { CK(cudaEventRecord(ev1, st)); c->prof_events.push_back(ev0); c->prof_events.push_back(ev1); }
        mark(st, "scan");
        {   // tail passes over the work-list of pixels the scan kernel left undecided (entry per lane; grid sized to the SM count)
            TailPassArgs TP{};
            TP.Wp = c->Wp; TP.WW = c->WW; TP.N = c->P.n_samples; TP.REQ = c->P.n_required; TP.plane = c->plane; TP.bg = c->bg;
            TP.wl_ctx = c->wl_ctx; TP.wl_cap = c->wl_cap; TP.hand = c->hand; TP.raw_bits = c->raw_alt; TP.lut = c->lut; TP.ctl = c->ctl;
            TP.collect_stats = c->collect_stats;
            cudaEvent_t tp0 = nullptr, tp1 = nullptr;
            if(c->profile) { CK(cudaEventCreate(&tp0)); CK(cudaEventCreate(&tp1)); CK(cudaEventRecord(tp0, st)); }
            // persistent grids (warps pull chunks of 32 entries): resident CTAs per SM x SM count, capped by the frame size
            const int npx = W * H;
            const int g1 = std::max(1, std::min(c->sm_count * TAIL1_MINB, (npx / 4 + 127) / 128)), g2 = std::max(1, std::min(c->sm_count * TAIL2_MINB, (npx / 8 + 127) / 128));
            TP.in_idx = nullptr; TP.in_count = &c->ctl->wl_count; TP.cursor = &c->ctl->wl_cursor; TP.out_idx = c->wl2_idx; TP.out_count = &c->ctl->wl2_count; TP.s_limit = TAIL_PASS1_LIMIT;
#if LVB_PDL
            // programmatic dependent launch: the pass becomes resident while its predecessor drains (the kernels order themselves with pdl_wait())
            cudaLaunchAttribute pdl_at[1];
            pdl_at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; pdl_at[0].val.programmaticStreamSerializationAllowed = 1;
            static const bool pdl_on = !(getenv("LVB_NO_PDL") && atoi(getenv("LVB_NO_PDL")));
            // frames up to 640x480 are bound by the host's launch rate, where the attribute costs more than the overlap returns (320x240: 57 -> 68 us)
            const bool pdl_size = (size_t)W * H > (size_t)640 * 480;
            auto pdl_launch = [&](void (*k)(const TailPassArgs), int g) {
                cudaLaunchConfig_t cfg{};
                cfg.gridDim = dim3((unsigned)g); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 0; cfg.stream = st; cfg.attrs = pdl_at; cfg.numAttrs = (pdl_on && pdl_size && !c->profile) ? 1 : 0;
                CK(cudaLaunchKernelEx(&cfg, k, TP));
            };
            if(c->lut_small) { if(C == 1) pdl_launch(subsense_tail_pass<1, true, TAIL1_B, TAIL1_MINB, false>, g1); else pdl_launch(subsense_tail_pass<3, true, TAIL1_B, TAIL1_MINB, false>, g1); }
            else { if(C == 1) pdl_launch(subsense_tail_pass<1, false, TAIL1_B, TAIL1_MINB, false>, g1); else pdl_launch(subsense_tail_pass<3, false, TAIL1_B, TAIL1_MINB, false>, g1); }
#else
            if(c->lut_small) { if(C == 1) subsense_tail_pass<1, true, TAIL1_B, TAIL1_MINB, false><<<g1, 128, 0, st>>>(TP); else subsense_tail_pass<3, true, TAIL1_B, TAIL1_MINB, false><<<g1, 128, 0, st>>>(TP); }
            else { if(C == 1) subsense_tail_pass<1, false, TAIL1_B, TAIL1_MINB, false><<<g1, 128, 0, st>>>(TP); else subsense_tail_pass<3, false, TAIL1_B, TAIL1_MINB, false><<<g1, 128, 0, st>>>(TP); }
#endif
            LAUNCHED();
            TP.in_idx = c->wl2_idx; TP.in_count = &c->ctl->wl2_count; TP.cursor = &c->ctl->wl2_cursor; TP.out_idx = nullptr; TP.out_count = nullptr; TP.s_limit = 0xFFFFFFFFu;
#if LVB_PDL
            if(c->lut_small) { if(C == 1) pdl_launch(subsense_tail_pass<1, true, TAIL2_B, TAIL2_MINB, true>, g2); else pdl_launch(subsense_tail_pass<3, true, TAIL2_B, TAIL2_MINB, true>, g2); }
            else { if(C == 1) pdl_launch(subsense_tail_pass<1, false, TAIL2_B, TAIL2_MINB, true>, g2); else pdl_launch(subsense_tail_pass<3, false, TAIL2_B, TAIL2_MINB, true>, g2); }
#else
            if(c->lut_small) { if(C == 1) subsense_tail_pass<1, true, TAIL2_B, TAIL2_MINB, true><<<g2, 128, 0, st>>>(TP); else subsense_tail_pass<3, true, TAIL2_B, TAIL2_MINB, true><<<g2, 128, 0, st>>>(TP); }
            else { if(C == 1) subsense_tail_pass<1, false, TAIL2_B, TAIL2_MINB, true><<<g2, 128, 0, st>>>(TP); else subsense_tail_pass<3, false, TAIL2_B, TAIL2_MINB, true><<<g2, 128, 0, st>>>(TP); }
#endif
            LAUNCHED();
            if(c->profile) { CK(cudaEventRecord(tp1, st)); c->prof3_events.push_back(tp0); c->prof3_events.push_back(tp1); }
            mark(st, "scan tail passes");
        }
        CK(cudaEventRecord(c->ev_scan, st));
        // ---- mask stream: chain of frame k, reading raw(k), writing blinks(k) / lastfg(k) / fin(k) into the spare buffers
        CK(cudaStreamWaitEvent(sp, c->ev_scan, 0));
        PostArgs P{};
        P.W = W; P.H = H; P.WW = c->WW; P.Wp = c->Wp; P.raw = c->raw_alt; P.lastraw = c->lastraw; P.lastrawblink = c->lastrawblink; P.blinks = c->blinks_alt;
        P.tmpA = c->tmpA; P.pre = c->pre; P.reach = c->reach; P.comb = c->comb; P.lastfg = c->lastfg_alt; P.dilinv = c->dilinv;
        P.out_mask = d_mask_out; P.out_pitch = (size_t)W; P.fin = c->fin; P.fin_in = c->fin; P.ctl = c->ctl; P.median_k = c->median_k;
        P.frame = c->sub_frame; P.avg_samples = c->P.n_samples_for_moving_avgs;
        pp_blink_close<<<wg, 256, 0, sp>>>(P); LAUNCHED(); mark(sp, "  post: blink_close");
        {
            HoleArgs Hh{};
            Hh.W = W; Hh.H = H; Hh.WW = c->WW; Hh.RS = c->uf_rs; Hh.pre = c->pre; Hh.raw = c->raw_alt; Hh.comb = c->comb;
            Hh.parent = c->uf_parent; Hh.rankbase = c->uf_rankbase;
            const int rb = (H + 7) / 8;
            pp_holes_init<<<rb, 256, 0, sp>>>(Hh); LAUNCHED();
            pp_holes_union<<<rb, 256, 0, sp>>>(Hh); LAUNCHED();
            pp_holes_combine<<<rb, 256, 0, sp>>>(Hh); LAUNCHED(); mark(sp, "  post: holes");
        }
        pp_median<<<mg, tb, 0, sp>>>(c->comb, c->lastfg_alt, d_mask_out, (size_t)W, W, H, c->WW, c->median_k); LAUNCHED(); mark(sp, "  post: median");
        chain_done_kernel<<<1, 1, 0, sp>>>(c->ctl, seq); LAUNCHED();
        if(mask_ready) CK(cudaEventRecord(mask_ready, sp));
        pp_dilate_blink<<<wg, 256, 0, sp>>>(P); LAUNCHED();
        mark(sp, "  post: dilate_blink");
        // the final-segmentation EMAs of this frame (:553-554) are folded in by the feedback kernel of the NEXT frame, which reads
        // them anyway (flush_pending() does it for anything that needs them earlier)
        // ---- instance stream: feedback(k) needs chain(k-1) (blinks / lastfg / fin of the previous frame) and the motion sum of frame k
        if(c->post_pending) CK(cudaStreamWaitEvent(st, c->ev_post, 0));
        CK(cudaStreamWaitEvent(st, c->ev_ds, 0));
        TailArgs T{};
        T.ctl = c->ctl; T.lut = c->lut; T.rel = c->P.rel_lbsp_threshold; T.lbsp_off = c->P.lbsp_threshold_offset; T.min_color = c->P.color_dist_threshold;
        T.avg_samples = c->P.n_samples_for_moving_avgs; T.N = c->P.n_samples; T.dsW = c->dsW; T.dsH = c->dsH; T.seed = c->seed; T.wait_seq = seq;
        cudaEvent_t fb0 = nullptr, fb1 = nullptr;
        if(c->profile) { CK(cudaEventCreate(&fb0)); CK(cudaEventCreate(&fb1)); CK(cudaEventRecord(fb0, st)); }
        // persistent: sm_count x FB_CTAS_PER_SM CTAs walk the 32x8 tiles (one tile each when the frame has fewer tiles than that)
        const int fb_tiles = (c->Wp / 32) * ((H + FB_H - 1) / FB_H);
        static const int fb_ctas = getenv("LVB_FB_CTAS") ? atoi(getenv("LVB_FB_CTAS")) : FB_GRID_CTAS_PER_SM;
        const dim3 fg((unsigned)std::min(fb_tiles, c->sm_count * fb_ctas)), fb(32, FB_H);
        if(C == 1) subsense_feedback<1><<<fg, fb, 0, st>>>(A, T); else subsense_feedback<3><<<fg, fb, 0, st>>>(A, T);
        LAUNCHED(); mark(st, "feedback");
        if(c->profile) { CK(cudaEventRecord(fb1, st)); c->prof2_events.push_back(fb0); c->prof2_events.push_back(fb1); }
        CK(cudaEventRecord(c->ev_post, sp)); c->post_pending = true; // recorded after feedback(k) is enqueued; covers the whole chain of frame k
        // the neighbour writes queued by feedback(k) ("phase B") are applied by scan(k+1), or by whoever needs the model first
        c->nb_seq = seq;
        std::swap(c->last_color, c->last_color_alt); std::swap(c->last_desc, c->last_desc_alt); c->scan_maps_idx ^= 1;
        c->ghost_idx ^= 1;
        std::swap(c->raw, c->raw_alt); std::swap(c->blinks, c->blinks_alt); std::swap(c->lastfg, c->lastfg_alt);
        c->fin_pending = c->sub_frame;
        c->sub_frame += 1;
        launch_refresh(c); mark(st, "refresh"); // reads lastfg(k), only when the frame tail requested it (the tail waited for the mask)
    } else { // LOBSTER
        if(c->lut_small) { if(C == 1) lobster_phaseA<1, true><<<stage_grid(c), stage_block, 0, st>>>(A, tmap); else lobster_phaseA<3, true><<<stage_grid(c), stage_block, 0, st>>>(A, tmap); }
        else { if(C == 1) lobster_phaseA<1, false><<<stage_grid(c), stage_block, 0, st>>>(A, tmap); else lobster_phaseA<3, false><<<stage_grid(c), stage_block, 0, st>>>(A, tmap); }
        LAUNCHED();
        if(c->profile) { CK(cudaEventRecord(ev1, st)); c->prof_events.push_back(ev0); c->prof_events.push_back(ev1); }
        B.bump_frame = &c->ctl->frame_idx; // the frame counter (Philox index) advances in phase B: no tail kernel
        // small frames are launch bound (BASELINE config #2, 320x240: 36 us of driver calls per frame with the fork / join below), so they
        // run phase B and the median back to back on the instance stream; large frames overlap the two on the auxiliary stream
        const bool serial = (size_t)W * H <= (size_t)640 * 480;
        cudaStream_t sb = serial ? st : c->s_aux;
        if(!serial) { CK(cudaEventRecord(c->ev_fork, st)); CK(cudaStreamWaitEvent(c->s_aux, c->ev_fork, 0)); }
        if(C == 1) neighbor_write_phaseB<1, 1><<<tg, tb, 0, sb>>>(B); else neighbor_write_phaseB<3, 1><<<tg, tb, 0, sb>>>(B);   // LOBSTER: 3x3 intents only
        LAUNCHED();
        if(!serial) CK(cudaEventRecord(c->ev_join, c->s_aux));
        pp_median<<<mg, tb, 0, st>>>(c->raw, c->lastfg, d_mask_out, (size_t)W, W, H, c->WW, c->median_k); LAUNCHED();
        if(!serial) CK(cudaStreamWaitEvent(st, c->ev_join, 0));
        if(mask_ready) CK(cudaEventRecord(mask_ready, st));
    }
    if(c->collect_stats) ++c->stat_frames;
}

void check_apply(lvb_context* c, const void* img, double lr) {
    REQUIRE(c->initialized, "algo & model must be initialized first");
    REQUIRE(img != nullptr, "input image type/size mismatch with initialization type/size");
    if(c->algo == LVB_ALGO_LOBSTER) REQUIRE(lr > 0, "learning rate must be a positive value; faster learning is achieved with smaller values");
    REQUIRE(!std::isnan(lr), "learning rate must not be NaN");
}

bool is_pinned(const void* p) {
    cudaPointerAttributes at;
    if(cudaPointerGetAttributes(&at, p) != cudaSuccess) { (void)cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost;
}
void apply_async(lvb_context* c, const uint8_t* img, uint8_t* mask, double lr) {
    check_apply(c, img, lr);
    REQUIRE(mask != nullptr, "output mask must be provided");
    CK(cudaSetDevice(c->device));
    lvb_context::Slot& sl = c->slot[c->n_submitted & 1];
    REQUIRE(!sl.busy, "two frames are already in flight: collect one with lvb_sync_next (or lvb_sync) first");
    const size_t row = (size_t)c->W * c->C;
    const uint8_t* src = img;
    if(!is_pinned(img)) { std::memcpy(sl.h_img, img, row * c->H); src = sl.h_img; } // pageable -> pinned staging
    // upload on s_in (overlaps the previous frame's kernels), kernels on the instance stream, mask back on s_out.
    // The slot's device frame may still be read by the frame that used it two submissions ago (its motion-analysis kernel runs on the
    // lowest-priority stream and is only awaited by that frame's feedback kernel, not by its mask): wait for "input consumed"
    if(sl.used) CK(cudaStreamWaitEvent(c->s_in, sl.in_consumed, 0));
    CK(cudaMemcpy2DAsync(sl.d_img, c->ipitch, src, row, row, c->H, cudaMemcpyHostToDevice, c->s_in));
    CK(cudaEventRecord(sl.h2d_done, c->s_in));
    CK(cudaStreamWaitEvent(c->stream, sl.h2d_done, 0));
    enqueue_frame(c, sl.d_img, c->ipitch, sl.tmap, sl.use_tma, sl.d_mask, lr, sl.compute_done);
    CK(cudaEventRecord(sl.in_consumed, c->stream)); sl.used = true; // behind every kernel of this frame that reads the input (the stream waited for the side streams' readers)
    CK(cudaStreamWaitEvent(c->s_out, sl.compute_done, 0));
    sl.direct = is_pinned(mask);
    CK(cudaMemcpyAsync(sl.direct ? mask : sl.h_mask, sl.d_mask, (size_t)c->W * c->H, cudaMemcpyDeviceToHost, c->s_out));
    CK(cudaEventRecord(sl.d2h_done, c->s_out));
    sl.user_mask = mask; sl.busy = true;
    ++c->n_submitted;
}
/// wait for the OLDEST frame in flight and deliver its mask; returns false when nothing is pending
bool sync_next(lvb_context* c) {
    if(c->n_collected == c->n_submitted) return false;
    lvb_context::Slot& sl = c->slot[c->n_collected & 1];
    CK(cudaEventSynchronize(sl.d2h_done));
    if(!sl.direct) std::memcpy(sl.user_mask, sl.h_mask, (size_t)c->W * c->H);
    sl.busy = false;
    ++c->n_collected;
    return true;
}
void sync(lvb_context* c) {
    while(sync_next(c)) {}
    sync_streams(c);
}

// ---- host-side enqueue pool: a frame is 12 kernel launches + ~10 event operations, so with many small streams the enqueueing
// thread is the bottleneck (BASELINE config #5: 64 VGA streams per GPU). The batch entry points spread the instances over a few
// persistent worker threads (the lv::WorkerPool pattern of apps/changedet/src/main.cpp:148-154); instances are independent.
class EnqueuePool {
public:
    EnqueuePool() {
        const char* e = getenv("LVB_ENQUEUE_THREADS");
        int n = e ? atoi(e) : (int)std::min(8u, std::max(1u, std::thread::hardware_concurrency() / 2));
        n = std::max(1, std::min(n, 64));
        for(int t = 1; t < n; ++t) workers_.emplace_back([this] { loop(); }); // the calling thread is worker 0
    }
    ~EnqueuePool() {
        { std::lock_guard<std::mutex> l(m_); stop_ = true; }
        cv_.notify_all();
        for(std::thread& t : workers_) t.join();
    }
    /// runs job(0..n-1), each index once, on the pool + the caller; rethrows the first failure
    void run(int n, const std::function<void(int)>& job) {
        if(n <= 0) return;
        // one batch at a time: concurrent callers (e.g. one host thread per GPU) queue up here instead of overwriting each other's
        // job descriptor while it is still being worked on
        std::lock_guard<std::mutex> batch(run_mutex_);
        std::unique_lock<std::mutex> l(m_);
        job_ = &job; n_ = n; next_ = 0; done_ = 0; err_.clear(); ++gen_;
        l.unlock();
        cv_.notify_all();
        work();
        l.lock();
        cv_done_.wait(l, [this] { return done_ == n_; });
        job_ = nullptr;
        if(!err_.empty()) throw std::runtime_error(err_);
    }
private:
    void work() {
        while(true) {
            int i;
            { std::lock_guard<std::mutex> l(m_); if(!job_ || next_ >= n_) return; i = next_++; }
            std::string err;
            try { (*job_)(i); } catch(const std::exception& e) { err = e.what(); }
            std::lock_guard<std::mutex> l(m_);
            if(!err.empty() && err_.empty()) err_ = err;
            if(++done_ == n_) cv_done_.notify_all();
        }
    }
    void loop() {
        uint64_t seen = 0;
        while(true) {
            { std::unique_lock<std::mutex> l(m_); cv_.wait(l, [&] { return stop_ || gen_ != seen; }); if(stop_) return; seen = gen_; }
            work();
        }
    }
    std::vector<std::thread> workers_;
    std::mutex m_, run_mutex_; std::condition_variable cv_, cv_done_;
    const std::function<void(int)>* job_ = nullptr;
    int n_ = 0, next_ = 0, done_ = 0; uint64_t gen_ = 0; bool stop_ = false; std::string err_;
};
EnqueuePool& enqueue_pool() { static EnqueuePool p; return p; }

// ---- state export / import in the reference's layout (tests + checkpointing) ----
struct StateDesc { const char* name; int kind; };
enum { K_BITS, K_MAPF, K_FIN, K_COLPLANE, K_DESCPLANE, K_BGCOL, K_BGDESC, K_LUT, K_DS, K_ROI, K_SCALARS };

uint32_t* bits_by_name(lvb_context* c, const std::string& n) {
    if(n == "lastfg") return c->lastfg; if(n == "unstable") return c->unstable; if(n == "blinks") return c->blinks;
    if(n == "lastraw") return c->lastraw; if(n == "lastrawblink") return c->lastrawblink; if(n == "dilinv") return c->dilinv;
    if(n == "rawmask") return c->raw; if(n == "ghost") return c->ghost[c->ghost_idx];
    if(c->algo == LVB_ALGO_PAWCS) { if(n == "illum") return c->illum; if(n == "dil") return c->dil; }
    return nullptr;
}
int map_index(const std::string& n) {
    static const char* names[8] = {"T", "R", "v", "Dlast", "DminLT", "DminST", "rawLT", "rawST"};
    for(int i = 0; i < 8; ++i) if(n == names[i]) return i;
    return -1;
}
size_t state_bytes(lvb_context* c, const std::string& n) {
    const size_t npx = (size_t)c->W * c->H;
    const bool paw = c->algo == LVB_ALGO_PAWCS;
    const bool sub = c->algo == LVB_ALGO_SUBSENSE || paw;
    if(n == "scalars") return 16 * sizeof(double);
    if(paw) {
        const size_t nw = (size_t)c->NW, ng = (size_t)c->NG, msz = (size_t)c->gW * c->gH;
        if(n == "lw_first" || n == "lw_last" || n == "lw_occ") return npx * nw * 4;
        if(n == "lw_color") return npx * nw * c->C;
        if(n == "lw_desc") return npx * nw * c->C * 2;
        if(n == "gw_weight") return ng * 4; if(n == "gw_map") return ng * msz * 4; if(n == "gw_bits") return ng;
        if(n == "gw_color") return ng * c->C; if(n == "gw_desc") return ng * c->C * 2; if(n == "gdict") return ng * 4; if(n == "glut") return npx * ng;
        if(n == "bg_color" || n == "bg_desc") throw std::runtime_error("unknown state buffer: " + n);
    }
    if(n == "roi") return npx;
    if(n == "lut") return 256;
    if(n == "lastcolor") return npx * c->C;
    if(n == "lastdesc") return npx * c->C * 2;
    if(n == "bg_color") return npx * c->C * c->P.n_samples;
    if(n == "bg_desc") return npx * c->C * 2 * c->P.n_samples;
    if(n == "lastfg" || n == "rawmask") return npx;
    if(sub) {
        if(bits_by_name(c, n)) return npx;
        if(map_index(n) >= 0 || n == "finLT" || n == "finST") return npx * 4;
        if(n == "dsLT" || n == "dsST") return (size_t)c->dsW * c->dsH * c->C * 4;
    }
    throw std::runtime_error("unknown state buffer: " + n);
}

void state_get(lvb_context* c, const std::string& n, void* out, size_t bytes) {
    REQUIRE(c->initialized, "algo must be initialized first");
    REQUIRE(bytes == state_bytes(c, n), "size mismatch for state buffer " + n);
    CK(cudaSetDevice(c->device));
    sync_streams(c);
    flush_pending(c);
    const int W = c->W, H = c->H, C = c->C, Wp = c->Wp, WW = c->WW;
    const size_t npx = (size_t)W * H;
    if(c->algo == LVB_ALGO_PAWCS) {
        const size_t nw = (size_t)c->NW, ng = (size_t)c->NG, msz = (size_t)c->gW * c->gH;
        if(n == "scalars") {
            FrameCtl f; get_ctl(c, f);
            GDict g; CK(cudaMemcpyAsync(&g, c->gd, sizeof(g), cudaMemcpyDeviceToHost, c->stream)); CK(cudaStreamSynchronize(c->stream));
            double* d = (double*)out; std::memset(d, 0, bytes);
            d[0] = (double)f.frame_idx - 1; d[1] = f.frames_since_reset; d[2] = f.cooldown; d[3] = f.auto_reset; d[4] = (double)nw; d[5] = (double)ng;
            d[6] = c->median_k; d[7] = g.weight_offset; d[8] = g.moving_camera; d[9] = g.last_nonflat_ratio; d[10] = (double)c->roi_count; d[11] = (double)c->orig_roi_count; d[12] = f.refresh_epoch;
            return;
        }
        if(n == "lw_first" || n == "lw_last" || n == "lw_occ" || n == "lw_color" || n == "lw_desc") {
            // device layout (pawcs.cuh): key = (occ, first + last); record = (colour, d0|d1<<16, d2, first) / 1 channel (colour | desc<<16, first)
            const size_t rw = c->paw_rec_bytes() / 4;
            std::vector<uint32_t> re(nw * c->plane * rw); std::vector<uint2> ke(nw * c->plane);
            d2h(c->stream, re.data(), c->lw_rec, re.size() * 4); d2h(c->stream, ke.data(), c->lw_key, ke.size() * 8);
            for(int y = 0; y < H; ++y) for(int x = 0; x < W; ++x) for(size_t i = 0; i < nw; ++i) {
                const size_t at = i * c->plane + (size_t)y * Wp + x, o = ((size_t)y * W + x) * nw + i;
                const uint32_t* r = &re[at * rw];
                const uint32_t first = r[rw - 1], last = ke[at].y - first;
                if(n == "lw_first") ((uint32_t*)out)[o] = (first == 1u && last == 0u) ? 0u : first; // "not created" marker (non-ROI pixels)
                else if(n == "lw_last") ((uint32_t*)out)[o] = last;
                else if(n == "lw_occ") ((uint32_t*)out)[o] = ke[at].x;
                else if(n == "lw_color") { for(int k = 0; k < C; ++k) ((uint8_t*)out)[o * C + k] = (uint8_t)(r[0] >> (8 * k)); }
                else { if(C == 1) ((uint16_t*)out)[o] = (uint16_t)(r[0] >> 16);
                       else { ((uint16_t*)out)[o * 3] = (uint16_t)(r[1] & 0xFFFFu); ((uint16_t*)out)[o * 3 + 1] = (uint16_t)(r[1] >> 16); ((uint16_t*)out)[o * 3 + 2] = (uint16_t)(r[2] & 0xFFFFu); } }
            }
            return;
        }
        if(n == "glut") {
            std::vector<uint8_t> h(ng * c->plane);
            d2h(c->stream, h.data(), c->glut, h.size());
            uint8_t* o = (uint8_t*)out;
            for(int y = 0; y < H; ++y) for(int x = 0; x < W; ++x) for(size_t i = 0; i < ng; ++i) o[((size_t)y * W + x) * ng + i] = h[i * c->plane + (size_t)y * Wp + x];
            return;
        }
        if(n == "gw_map") { d2h(c->stream, out, c->gmap, ng * msz * 4); return; }
        if(n == "gw_weight" || n == "gw_bits" || n == "gw_color" || n == "gw_desc" || n == "gdict") {
            GDict g; CK(cudaMemcpyAsync(&g, c->gd, sizeof(g), cudaMemcpyDeviceToHost, c->stream)); CK(cudaStreamSynchronize(c->stream));
            for(size_t i = 0; i < ng; ++i) {
                if(n == "gw_weight") ((float*)out)[i] = g.weight[i];
                else if(n == "gw_bits") ((uint8_t*)out)[i] = (uint8_t)g.bits[i];
                else if(n == "gdict") ((int32_t*)out)[i] = g.dict[i];
                else if(n == "gw_color") for(int k = 0; k < C; ++k) ((uint8_t*)out)[i * C + k] = (uint8_t)(g.color[i] >> (8 * k));
                else for(int k = 0; k < C; ++k) ((uint16_t*)out)[i * C + k] = (uint16_t)(k == 0 ? g.desc[i].x & 0xFFFFu : k == 1 ? g.desc[i].x >> 16 : g.desc[i].y & 0xFFFFu);
            }
            return;
        }
    }
    if(n == "scalars") {
        FrameCtl f; get_ctl(c, f);
        double* d = (double*)out; std::memset(d, 0, bytes);
        d[0] = (double)f.frame_idx - 1; d[1] = f.frames_since_reset; d[2] = f.cooldown; d[3] = f.auto_reset; d[4] = f.lr_scaling; d[5] = f.use3x3;
        d[6] = c->median_k; d[7] = f.t_lower; d[8] = f.t_upper; d[9] = f.last_nonzero_ratio; d[10] = (double)c->roi_count; d[11] = (double)c->orig_roi_count; d[12] = f.refresh_epoch;
        return;
    }
    if(n == "roi") { std::memcpy(out, c->roi_host.data(), npx); return; }
    if(n == "lut") { d2h(c->stream, out, c->lut, 256); return; }
    if(uint32_t* b = bits_by_name(c, n)) {
        std::vector<uint32_t> h((size_t)H * WW);
        d2h(c->stream, h.data(), b, h.size() * 4);
        uint8_t* o = (uint8_t*)out;
        const bool as01 = (n == "unstable" || n == "illum");
        for(int y = 0; y < H; ++y) for(int x = 0; x < W; ++x) { const bool v = (h[(size_t)y * WW + (x >> 5)] >> (x & 31)) & 1u; o[(size_t)y * W + x] = v ? (as01 ? 1 : 255) : 0; }
        return;
    }
    const int mi = map_index(n);
    if(mi >= 0) {
        std::vector<float> h(c->plane * 8);
        d2h(c->stream, h.data(), c->maps, h.size() * 4);
        float* o = (float*)out;
        for(int y = 0; y < H; ++y) for(int x = 0; x < W; ++x) o[(size_t)y * W + x] = h[((size_t)y * Wp + x) * 8 + mi];
        return;
    }
    if(n == "finLT" || n == "finST") {
        std::vector<float> h(c->plane * 2);
        d2h(c->stream, h.data(), c->fin, h.size() * 4);
        float* o = (float*)out; const int k = n == "finLT" ? 0 : 1;
        for(int y = 0; y < H; ++y) for(int x = 0; x < W; ++x) o[(size_t)y * W + x] = h[((size_t)y * Wp + x) * 2 + k];
        return;
    }
    if(n == "dsLT" || n == "dsST") { d2h(c->stream, out, n == "dsLT" ? c->dsLT : c->dsST, bytes); return; }
    auto unpack_col = [&](const uint8_t* h, uint8_t* o) {
        for(int y = 0; y < H; ++y) for(int x = 0; x < W; ++x) for(int k = 0; k < C; ++k)
            o[((size_t)y * W + x) * C + k] = C == 1 ? h[(size_t)y * Wp + x] : h[((size_t)y * Wp + x) * 4 + k];
    };
    auto unpack_desc = [&](const uint16_t* h, uint16_t* o) {
        for(int y = 0; y < H; ++y) for(int x = 0; x < W; ++x) for(int k = 0; k < C; ++k)
            o[((size_t)y * W + x) * C + k] = C == 1 ? h[(size_t)y * Wp + x] : h[((size_t)y * Wp + x) * 4 + k];
    };
    if(n == "lastcolor") { std::vector<uint8_t> h(c->plane * c->col_bytes()); d2h(c->stream, h.data(), c->last_color, h.size()); unpack_col(h.data(), (uint8_t*)out); return; }
    if(n == "lastdesc") { std::vector<uint16_t> h(c->plane * c->desc_bytes() / 2); d2h(c->stream, h.data(), c->last_desc, h.size() * 2); unpack_desc(h.data(), (uint16_t*)out); return; }
    if(n == "bg_color" || n == "bg_desc") { // one plane of sample records at a time: colour = bytes 0..C-1, descriptors = u16 at byte 4 (3ch) / 2 (1ch)
        const size_t rb = c->rec_bytes(), doff = C == 1 ? 2 : 4;
        std::vector<uint8_t> h(c->plane * rb);
        for(int s = 0; s < c->P.n_samples; ++s) {
            d2h(c->stream, h.data(), (uint8_t*)c->bg + (size_t)s * h.size(), h.size());
            for(int y = 0; y < H; ++y) for(int x = 0; x < W; ++x) for(int k = 0; k < C; ++k) {
                const uint8_t* r = h.data() + ((size_t)y * Wp + x) * rb;
                const size_t o = (size_t)s * npx * C + ((size_t)y * W + x) * C + k;
                if(n == "bg_color") ((uint8_t*)out)[o] = r[k];
                else { uint16_t v; std::memcpy(&v, r + doff + 2 * k, 2); ((uint16_t*)out)[o] = v; }
            }
        }
        return;
    }
    throw std::runtime_error("unknown state buffer: " + n);
}

void state_set(lvb_context* c, const std::string& n, const void* in, size_t bytes) {
    REQUIRE(c->initialized, "algo must be initialized first");
    REQUIRE(bytes == state_bytes(c, n), "size mismatch for state buffer " + n);
    CK(cudaSetDevice(c->device));
    sync_streams(c);
    flush_pending(c);
    const int W = c->W, H = c->H, C = c->C, Wp = c->Wp, WW = c->WW;
    const size_t npx = (size_t)W * H;
    if(c->algo == LVB_ALGO_PAWCS) {
        const size_t nw = (size_t)c->NW, ng = (size_t)c->NG, msz = (size_t)c->gW * c->gH;
        if(n == "scalars") {
            const double* d = (const double*)in;
            FrameCtl f; get_ctl(c, f);
            GDict g; CK(cudaMemcpyAsync(&g, c->gd, sizeof(g), cudaMemcpyDeviceToHost, c->stream)); CK(cudaStreamSynchronize(c->stream));
            f.frame_idx = (uint32_t)d[0] + 1; f.frames_since_reset = (uint32_t)d[1]; f.cooldown = (uint32_t)d[2]; f.auto_reset = d[3] != 0; f.refresh_epoch = (uint32_t)d[12];
            g.weight_offset = (uint32_t)d[7]; g.moving_camera = d[8] != 0; g.last_nonflat_ratio = (float)d[9];
            const bool boot = f.frame_idx <= PAW_BOOTSTRAP;
            const uint32_t avg = (uint32_t)c->P.n_samples_for_moving_avgs, nLT = boot ? avg / 2u : avg, nST = nLT / 4u;
            g.boot = boot; g.nST = nST;
            f.aLT = 1.0f / (float)std::min(f.frame_idx, nLT); f.aST = 1.0f / (float)std::min(f.frame_idx, nST);
            put_ctl(c, f);
            CK(cudaMemcpyAsync(c->gd, &g, sizeof(g), cudaMemcpyHostToDevice, c->stream)); CK(cudaStreamSynchronize(c->stream));
            c->paw_frame = f.frame_idx;
            return;
        }
        if(n == "lw_first" || n == "lw_last" || n == "lw_occ" || n == "lw_color" || n == "lw_desc") {
            const size_t rw = c->paw_rec_bytes() / 4;
            std::vector<uint32_t> re(nw * c->plane * rw); std::vector<uint2> ke(nw * c->plane);
            d2h(c->stream, re.data(), c->lw_rec, re.size() * 4); d2h(c->stream, ke.data(), c->lw_key, ke.size() * 8);
            for(int y = 0; y < H; ++y) for(int x = 0; x < W; ++x) for(size_t i = 0; i < nw; ++i) {
                const size_t at = i * c->plane + (size_t)y * Wp + x, o = ((size_t)y * W + x) * nw + i;
                uint32_t* r = &re[at * rw];
                const uint32_t first = r[rw - 1], last = ke[at].y - first;
                if(n == "lw_first") { const uint32_t v = ((const uint32_t*)in)[o]; r[rw - 1] = v; ke[at].y = v + last; }
                else if(n == "lw_last") ke[at].y = first + ((const uint32_t*)in)[o];
                else if(n == "lw_occ") ke[at].x = ((const uint32_t*)in)[o];
                else if(n == "lw_color") {
                    uint32_t v = 0; for(int k = 0; k < C; ++k) v |= (uint32_t)((const uint8_t*)in)[o * C + k] << (8 * k);
                    r[0] = C == 1 ? ((r[0] & 0xFFFF0000u) | v) : v;
                } else {
                    const uint16_t* d = (const uint16_t*)in + o * C;
                    if(C == 1) r[0] = (r[0] & 0x0000FFFFu) | ((uint32_t)d[0] << 16);
                    else { r[1] = (uint32_t)d[0] | ((uint32_t)d[1] << 16); r[2] = d[2]; }
                }
            }
            h2d(c->stream, c->lw_rec, re.data(), re.size() * 4); h2d(c->stream, c->lw_key, ke.data(), ke.size() * 8);
            return;
        }
        if(n == "glut") {
            std::vector<uint8_t> h(ng * c->plane, 0);
            const uint8_t* sI = (const uint8_t*)in;
            for(int y = 0; y < H; ++y) for(int x = 0; x < W; ++x) for(size_t i = 0; i < ng; ++i) h[i * c->plane + (size_t)y * Wp + x] = sI[((size_t)y * W + x) * ng + i];
            h2d(c->stream, c->glut, h.data(), h.size());
            return;
        }
        if(n == "gw_map") { h2d(c->stream, c->gmap, in, ng * msz * 4); return; }
        if(n == "gw_weight" || n == "gw_bits" || n == "gw_color" || n == "gw_desc" || n == "gdict") {
            GDict g; CK(cudaMemcpyAsync(&g, c->gd, sizeof(g), cudaMemcpyDeviceToHost, c->stream)); CK(cudaStreamSynchronize(c->stream));
            for(size_t i = 0; i < ng; ++i) {
                if(n == "gw_weight") g.weight[i] = ((const float*)in)[i];
                else if(n == "gw_bits") g.bits[i] = ((const uint8_t*)in)[i];
                else if(n == "gdict") g.dict[i] = ((const int32_t*)in)[i];
                else if(n == "gw_color") { uint32_t v = 0; for(int k = 0; k < C; ++k) v |= (uint32_t)((const uint8_t*)in)[i * C + k] << (8 * k); g.color[i] = v; }
                else { const uint16_t* dd = (const uint16_t*)in + i * C; g.desc[i] = C == 1 ? make_uint2(dd[0], 0) : make_uint2((uint32_t)dd[0] | ((uint32_t)dd[1] << 16), dd[2]); }
            }
            CK(cudaMemcpyAsync(c->gd, &g, sizeof(g), cudaMemcpyHostToDevice, c->stream)); CK(cudaStreamSynchronize(c->stream));
            return;
        }
    }
    if(n == "scalars") {
        const double* d = (const double*)in;
        FrameCtl f; get_ctl(c, f);
        f.frame_idx = (uint32_t)d[0] + 1; f.frames_since_reset = (uint32_t)d[1]; f.cooldown = (uint32_t)d[2]; f.auto_reset = d[3] != 0;
        f.refresh_epoch = (uint32_t)d[12];
        if(c->algo == LVB_ALGO_SUBSENSE) {
            f.lr_scaling = d[4] != 0; f.use3x3 = d[5] != 0; c->median_k = (int)d[6]; f.median_k = c->median_k;
            f.t_lower = (float)d[7]; f.t_upper = (float)d[8]; f.last_nonzero_ratio = (float)d[9];
            const uint32_t avg = (uint32_t)c->P.n_samples_for_moving_avgs;
            f.aLT = 1.0f / (float)std::min(f.frame_idx, avg); f.aST = 1.0f / (float)std::min(f.frame_idx, avg / 4u);
            c->sub_frame = f.frame_idx;
        }
        put_ctl(c, f);
        return;
    }
    REQUIRE(n != "roi", "use lvb_set_roi");
    if(n == "lut") { h2d(c->stream, c->lut, in, 256); c->lut_small = lut_fits_7bit(c, (const uint8_t*)in); return; }
    if(uint32_t* b = bits_by_name(c, n)) {
        std::vector<uint32_t> h((size_t)H * WW, 0);
        const uint8_t* s = (const uint8_t*)in;
        for(int y = 0; y < H; ++y) for(int x = 0; x < W; ++x) if(s[(size_t)y * W + x]) h[(size_t)y * WW + (x >> 5)] |= 1u << (x & 31);
        h2d(c->stream, b, h.data(), h.size() * 4);
        return;
    }
    const int mi = map_index(n);
    if(mi >= 0) {
        std::vector<float> h(c->plane * 8);
        d2h(c->stream, h.data(), c->maps, h.size() * 4);
        const float* s = (const float*)in;
        for(int y = 0; y < H; ++y) for(int x = 0; x < W; ++x) h[((size_t)y * Wp + x) * 8 + mi] = s[(size_t)y * W + x];
        h2d(c->stream, c->maps, h.data(), h.size() * 4);
        if(mi == 1 && c->r_plane) { // the scan kernel reads R(x) from its compact copy
            std::vector<float> rp(c->plane, 1.0f);
            for(int y = 0; y < H; ++y) for(int x = 0; x < W; ++x) rp[(size_t)y * Wp + x] = s[(size_t)y * W + x];
            h2d(c->stream, c->r_plane, rp.data(), rp.size() * 4);
        }
        if(mi == 3 || mi == 7) { // ghost flag is derived state: rawST > 0.995 && Dlast < 0.01 inside the ROI
            std::vector<uint32_t> g((size_t)H * WW, 0);
            for(int y = 0; y < H; ++y) for(int x = 0; x < W; ++x) {
                const float* px = &h[((size_t)y * Wp + x) * 8];
                if(c->roi_host[(size_t)y * W + x] && px[7] > 0.995f && px[3] < 0.010f) g[(size_t)y * WW + (x >> 5)] |= 1u << (x & 31);
            }
            h2d(c->stream, c->ghost[c->ghost_idx], g.data(), g.size() * 4);
        }
        return;
    }
    if(n == "finLT" || n == "finST") {
        std::vector<float> h(c->plane * 2);
        d2h(c->stream, h.data(), c->fin, h.size() * 4);
        const float* s = (const float*)in; const int k = n == "finLT" ? 0 : 1;
        for(int y = 0; y < H; ++y) for(int x = 0; x < W; ++x) h[((size_t)y * Wp + x) * 2 + k] = s[(size_t)y * W + x];
        h2d(c->stream, c->fin, h.data(), h.size() * 4);
        return;
    }
    if(n == "dsLT" || n == "dsST") { h2d(c->stream, n == "dsLT" ? c->dsLT : c->dsST, in, bytes); return; }
    auto pack_col = [&](const uint8_t* s, uint8_t* h) {
        for(int y = 0; y < H; ++y) for(int x = 0; x < W; ++x) for(int k = 0; k < C; ++k) {
            if(C == 1) h[(size_t)y * Wp + x] = s[(size_t)y * W + x]; else h[((size_t)y * Wp + x) * 4 + k] = s[((size_t)y * W + x) * C + k];
        }
    };
    auto pack_desc = [&](const uint16_t* s, uint16_t* h) {
        for(int y = 0; y < H; ++y) for(int x = 0; x < W; ++x) for(int k = 0; k < C; ++k) {
            if(C == 1) h[(size_t)y * Wp + x] = s[(size_t)y * W + x]; else h[((size_t)y * Wp + x) * 4 + k] = s[((size_t)y * W + x) * C + k];
        }
    };
    if(n == "lastcolor") { std::vector<uint8_t> h(c->plane * c->col_bytes(), 0); pack_col((const uint8_t*)in, h.data()); h2d(c->stream, c->last_color, h.data(), h.size());
                           if(c->last_color_alt) h2d(c->stream, c->last_color_alt, h.data(), h.size()); return; }
    if(n == "lastdesc") { std::vector<uint16_t> h(c->plane * c->desc_bytes() / 2, 0); pack_desc((const uint16_t*)in, h.data()); h2d(c->stream, c->last_desc, h.data(), h.size() * 2);
                          if(c->last_desc_alt) h2d(c->stream, c->last_desc_alt, h.data(), h.size() * 2); return; }
    if(n == "bg_color" || n == "bg_desc") { // read-modify-write of the record planes (the other half of each record is kept)
        const size_t rb = c->rec_bytes(), doff = C == 1 ? 2 : 4;
        std::vector<uint8_t> h(c->plane * rb);
        for(int s = 0; s < c->P.n_samples; ++s) {
            d2h(c->stream, h.data(), (uint8_t*)c->bg + (size_t)s * h.size(), h.size());
            for(int y = 0; y < H; ++y) for(int x = 0; x < W; ++x) for(int k = 0; k < C; ++k) {
                uint8_t* r = h.data() + ((size_t)y * Wp + x) * rb;
                const size_t o = (size_t)s * npx * C + ((size_t)y * W + x) * C + k;
                if(n == "bg_color") r[k] = ((const uint8_t*)in)[o];
                else { const uint16_t v = ((const uint16_t*)in)[o]; std::memcpy(r + doff + 2 * k, &v, 2); }
            }
            h2d(c->stream, (uint8_t*)c->bg + (size_t)s * h.size(), h.data(), h.size());
        }
        if(n == "bg_color") rebuild_cbox(c);
        return;
    }
    throw std::runtime_error("unknown state buffer: " + n);
}

void get_bg_image(lvb_context* c, uint8_t* out_color, uint16_t* out_desc, bool out_on_device = false) {
    REQUIRE(c->initialized, "algo must be initialized first");
    CK(cudaSetDevice(c->device));
    sync_streams(c);
    flush_pending(c);
    const size_t n = (size_t)c->W * c->H * c->C;
    uint8_t* dc = nullptr; uint16_t* dd = nullptr;
    if(out_color) dc = dalloc<uint8_t>(c->stream, n, false); else dd = dalloc<uint16_t>(c->stream, n, false);
    const cudaMemcpyKind kind = out_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    if(c->algo == LVB_ALGO_PAWCS) {
        const PawArgs A = paw_args(c, c->d_img, c->ipitch, c->use_tma, 0.0);
        if(c->C == 1) pawcs_background_kernel<1><<<tile_grid(c), dim3(32, 8), 0, c->stream>>>(A, dc, dd, 1); else pawcs_background_kernel<3><<<tile_grid(c), dim3(32, 8), 0, c->stream>>>(A, dc, dd, 1);
    }
    else if(c->C == 1) background_image_kernel<1><<<tile_grid(c), dim3(32, 8), 0, c->stream>>>(c->bg, c->plane, c->P.n_samples, c->W, c->H, c->Wp, dc, dd);
    else background_image_kernel<3><<<tile_grid(c), dim3(32, 8), 0, c->stream>>>(c->bg, c->plane, c->P.n_samples, c->W, c->H, c->Wp, dc, dd);
    ++g_launches;
    cudaError_t e = cudaGetLastError();
    if(e == cudaSuccess) e = out_color ? cudaMemcpyAsync(out_color, dc, n, kind, c->stream) : cudaMemcpyAsync(out_desc, dd, n * 2, kind, c->stream);
    if(e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(dc); cudaFree(dd);
    CK(e);
}

} // namespace

#define LVB_TRY try {
#define LVB_CATCH } catch(const std::exception& e) { g_err = e.what(); return 1; } return 0;

extern "C" {

const char* lvb_last_error(void) { return g_err.c_str(); }

int lvb_device_count(void) {
    int n = 0;
    if(cudaGetDeviceCount(&n) != cudaSuccess) { (void)cudaGetLastError(); return 0; }
    return n;
}

int lvb_default_params(int algo, lvb_params* out) {
    LVB_TRY
    REQUIRE(out != nullptr, "null output");
    REQUIRE(algo >= 0 && algo <= 2, "unknown algorithm id");
    lvb_params p{};
    p.rel_lbsp_threshold = 0.333f; p.lbsp_threshold_offset = 0; p.median_blur_kernel_size = 9; p.n_samples_for_moving_avgs = 100; p.n_global_words = 25;
    if(algo == LVB_ALGO_LOBSTER) { p.desc_dist_threshold = 4; p.color_dist_threshold = 30; p.n_samples = 35; p.n_required = 2; }
    if(algo == LVB_ALGO_SUBSENSE) { p.desc_dist_threshold = 3; p.color_dist_threshold = 30; p.n_samples = 50; p.n_required = 2; }
    if(algo == LVB_ALGO_PAWCS) { p.desc_dist_threshold = 2; p.color_dist_threshold = 20; p.n_samples = 50; p.n_required = 0; }
    *out = p;
    LVB_CATCH
}
double lvb_default_learning_rate(int algo) { return algo == LVB_ALGO_LOBSTER ? 16.0 : 0.0; }

/// per-device settings applied before the first instance is created on a device.
/// L2 fetch granularity: the scan tail, the stochastic sample writes (read-modify-write of a 32-byte sector) and the queued
/// neighbour writes touch lone 32-byte sectors; the default 64-byte fetch granularity doubles their DRAM traffic.
static void device_tuning_once(int device) {
    static std::mutex mu; static std::vector<int> done;
    std::lock_guard<std::mutex> lk(mu);
    if(std::find(done.begin(), done.end(), device) != done.end()) return;
    done.push_back(device);
    size_t gran = 32;
    if(const char* e = getenv("LVB_L2_FETCH")) gran = (size_t)atoi(e);
    if(gran == 32 || gran == 64 || gran == 128) { if(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran) != cudaSuccess) (void)cudaGetLastError(); }
}

int lvb_create(int algo, const lvb_params* params, int device, uint64_t seed, lvb_handle* out) {
    LVB_TRY
    REQUIRE(out != nullptr, "null output handle");
    REQUIRE(algo == LVB_ALGO_LOBSTER || algo == LVB_ALGO_SUBSENSE || algo == LVB_ALGO_PAWCS, "unknown algorithm id");
    lvb_params p;
    if(params) p = *params; else lvb_default_params(algo, &p);
    REQUIRE(p.n_samples > 0 && p.n_required <= p.n_samples, "algo cannot require more sample matches than sample count in model");
    REQUIRE(p.n_samples <= 255, "at most 255 samples per pixel are supported");
    REQUIRE(p.color_dist_threshold > 0 && p.desc_dist_threshold > 0, "distance thresholds must be positive values");
    REQUIRE(algo == LVB_ALGO_PAWCS ? p.n_required >= 0 : p.n_required >= 1, "the number of required sample matches must be positive");
    REQUIRE(p.rel_lbsp_threshold >= 0, "relative threshold for LBSP features must be non-negative");
    REQUIRE(p.n_samples_for_moving_avgs >= 4, "moving average window must be >= 4");
    int ndev = lvb_device_count();
    REQUIRE(ndev > 0, "no CUDA device available: litiv_b200 has no CPU fallback");
    REQUIRE(device >= 0 && device < ndev, "invalid CUDA device id");
    CK(cudaSetDevice(device));
    device_tuning_once(device);
    lvb_context* c = new lvb_context();
    c->algo = algo; c->device = device; c->seed = seed;
    if(cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || c->sm_count <= 0) { (void)cudaGetLastError(); c->sm_count = 148; }
    static_assert(sizeof(Params) == sizeof(lvb_params), "params mirror out of sync");
    std::memcpy(&c->P, &p, sizeof(p));
    // high priority: while phase B occupies the auxiliary (low-priority) stream the small mask kernels get SM slots first
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    // three priority levels per instance: mask stream (highest: its small kernels must slip in beside the big ones) > instance
    // stream > auxiliary stream (lowest)
    const int prio_mid = prio_hi < prio_lo - 1 ? prio_hi + 1 : prio_hi;
    cudaError_t e = cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, prio_mid);
    if(e != cudaSuccess) { delete c; CK(e); }
    *out = c;
    LVB_CATCH
}
int lvb_destroy(lvb_handle h) {
    if(!h) return 0;
    cudaSetDevice(h->device);
    if(h->stream) { cudaStreamSynchronize(h->stream); }
    if(h->s_post) cudaStreamSynchronize(h->s_post);
    if(h->s_aux) cudaStreamSynchronize(h->s_aux);
    h->free_all();
    for(auto& sl : h->slot) { if(sl.in_consumed) cudaEventDestroy(sl.in_consumed); if(sl.h2d_done) cudaEventDestroy(sl.h2d_done); if(sl.compute_done) cudaEventDestroy(sl.compute_done); if(sl.d2h_done) cudaEventDestroy(sl.d2h_done); }
    if(h->s_aux) cudaStreamDestroy(h->s_aux);
    if(h->s_post) cudaStreamDestroy(h->s_post);
    for(cudaEvent_t e : {h->ev_scan, h->ev_post, h->ev_ds, h->ev_mask}) if(e) cudaEventDestroy(e);
    if(h->ev_fork) cudaEventDestroy(h->ev_fork);
    if(h->ev_join) cudaEventDestroy(h->ev_join);
    if(h->s_in) cudaStreamDestroy(h->s_in);
    if(h->s_out) cudaStreamDestroy(h->s_out);
    if(h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return 0;
}
void* lvb_stream(lvb_handle h) { return h ? (void*)h->stream : nullptr; }

int lvb_initialize(lvb_handle h, const uint8_t* img, int width, int height, int channels, size_t step, const uint8_t* roi) {
    LVB_TRY
    REQUIRE(h != nullptr, "null handle");
    do_initialize(h, img, width, height, channels, step, roi);
    LVB_CATCH
}
int lvb_apply(lvb_handle h, const uint8_t* img, uint8_t* fgmask, double lr) {
    LVB_TRY
    REQUIRE(h != nullptr, "null handle");
    apply_async(h, img, fgmask, lr);
    // returns as soon as this frame's mask has landed in the caller's buffer: the feedback kernel and the rest of the mask chain
    // keep running on the instance's streams while the caller fetches / uploads the next frame (anything that reads the model
    // or the state waits for them)
    while(sync_next(h)) {}
    LVB_CATCH
}
int lvb_apply_async(lvb_handle h, const uint8_t* img, uint8_t* fgmask, double lr) {
    LVB_TRY
    REQUIRE(h != nullptr, "null handle");
    apply_async(h, img, fgmask, lr);
    LVB_CATCH
}
int lvb_apply_stream(lvb_handle h, const uint8_t* const* imgs, uint8_t* const* fgmasks, int n, const double* lrs) {
    LVB_TRY
    REQUIRE(h != nullptr && imgs && fgmasks && lrs && n >= 0, "bad stream arguments");
    CK(cudaSetDevice(h->device));
    for(int i = 0; i < n; ++i) {
        apply_async(h, imgs[i], fgmasks[i], lrs[i]);   // upload of frame i overlaps the kernels of frame i-1
        if(i > 0) REQUIRE(sync_next(h), "no frame in flight");
    }
    if(n > 0) REQUIRE(sync_next(h), "no frame in flight");
    LVB_CATCH
}
int lvb_sync(lvb_handle h) {
    LVB_TRY
    REQUIRE(h != nullptr, "null handle");
    CK(cudaSetDevice(h->device));
    sync(h);
    if(h->initialized && h->algo == LVB_ALGO_SUBSENSE) { // a model reset that gave up waiting for its final mask is an error, not a silent skip
        FrameCtl f; get_ctl(h, f);
        if(f.spin_timeout) { const uint32_t seq = f.spin_timeout; f.spin_timeout = 0; put_ctl(h, f);
            throw std::runtime_error("the mask chain of frame " + std::to_string(seq) + " did not complete within the wait bound of the frame-level model reset (streams not progressing concurrently?); the reset was skipped"); }
    }
    LVB_CATCH
}
int lvb_sync_next(lvb_handle h) {
    LVB_TRY
    REQUIRE(h != nullptr, "null handle");
    CK(cudaSetDevice(h->device));
    REQUIRE(sync_next(h), "no frame in flight");
    LVB_CATCH
}
int lvb_apply_batch(lvb_handle* hs, const uint8_t* const* imgs, uint8_t* const* masks, int n, double lr) {
    LVB_TRY
    REQUIRE(hs && imgs && masks && n >= 0, "bad batch arguments");
    for(int i = 0; i < n; ++i) REQUIRE(hs[i] != nullptr, "null handle in batch");
    enqueue_pool().run(n, [&](int i) { apply_async(hs[i], imgs[i], masks[i], lr); });
    for(int i = 0; i < n; ++i) { CK(cudaSetDevice(hs[i]->device)); while(sync_next(hs[i])) {} }
    LVB_CATCH
}
int lvb_apply_batch_device(lvb_handle* hs, const uint8_t* const* d_imgs, size_t d_step, uint8_t* const* d_masks, int n, double lr) {
    LVB_TRY
    REQUIRE(hs && d_imgs && n >= 0, "bad batch arguments");
    for(int i = 0; i < n; ++i) {
        REQUIRE(hs[i] != nullptr, "null handle in batch");
        check_apply(hs[i], d_imgs[i], lr);
        REQUIRE(d_step >= (size_t)hs[i]->W * hs[i]->C, "row step smaller than a row");
    }
    enqueue_pool().run(n, [&](int i) {
        lvb_context* h = hs[i];
        CK(cudaSetDevice(h->device));
        if(h->ext_ptr != d_imgs[i] || h->ext_pitch != d_step) {
            h->ext_tma = make_image_tmap(&h->tmap_ext, d_imgs[i], h->W, h->H, h->C, d_step) ? 1 : 0;
            h->ext_ptr = d_imgs[i]; h->ext_pitch = d_step;
        }
        enqueue_frame(h, d_imgs[i], d_step, h->tmap_ext, h->ext_tma, (d_masks && d_masks[i]) ? d_masks[i] : h->d_mask, lr);
    });
    LVB_CATCH
}
int lvb_apply_device(lvb_handle h, const uint8_t* d_img, size_t d_step, uint8_t* d_mask, double lr) {
    LVB_TRY
    REQUIRE(h != nullptr, "null handle");
    check_apply(h, d_img, lr);
    REQUIRE(d_step >= (size_t)h->W * h->C, "row step smaller than a row");
    CK(cudaSetDevice(h->device));
    if(h->ext_ptr != d_img || h->ext_pitch != d_step) {
        h->ext_tma = make_image_tmap(&h->tmap_ext, d_img, h->W, h->H, h->C, d_step) ? 1 : 0;
        h->ext_ptr = d_img; h->ext_pitch = d_step;
    }
    enqueue_frame(h, d_img, d_step, h->tmap_ext, h->ext_tma, d_mask ? d_mask : h->d_mask, lr);
    LVB_CATCH
}
int lvb_flush(lvb_handle h) {
    LVB_TRY
    REQUIRE(h != nullptr, "null handle");
    CK(cudaSetDevice(h->device));
    if(h->post_pending) CK(cudaStreamWaitEvent(h->stream, h->ev_post, 0));
    LVB_CATCH
}
int lvb_get_background_image(lvb_handle h, uint8_t* out) {
    LVB_TRY
    REQUIRE(h && out, "null argument");
    get_bg_image(h, out, nullptr);
    LVB_CATCH
}
int lvb_get_background_image_device(lvb_handle h, uint8_t* d_out) {
    LVB_TRY
    REQUIRE(h && d_out, "null argument");
    cudaPointerAttributes at;
    REQUIRE(cudaPointerGetAttributes(&at, d_out) == cudaSuccess && (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged), "output must be a device pointer");
    get_bg_image(h, d_out, nullptr, true);
    LVB_CATCH
}
int lvb_validate_roi(uint8_t* roi, int width, int height, int border) {
    LVB_TRY
    // IIBackgroundSubtractor::validateROI (BackgroundSubtractionUtils.cpp:28-36): everything within `border` pixels of the frame edge is cleared
    REQUIRE(roi != nullptr && width > 0 && height > 0, "provided ROI must be non-empty and of type 8UC1");
    REQUIRE(border >= 0, "border size must be non-negative");
    for(int y = 0; y < height; ++y) for(int x = 0; x < width; ++x)
        if(x < border || y < border || x >= width - border || y >= height - border) roi[(size_t)y * width + x] = 0;
    LVB_CATCH
}
int lvb_get_background_descriptors_image(lvb_handle h, uint16_t* out) {
    LVB_TRY
    REQUIRE(h && out, "null argument");
    get_bg_image(h, nullptr, out);
    LVB_CATCH
}
int lvb_refresh_model(lvb_handle h, float frac, int force_fg) {
    LVB_TRY
    REQUIRE(h != nullptr, "null handle");
    REQUIRE(h->algo != LVB_ALGO_PAWCS, "PAWCS: use lvb_pawcs_refresh_model(base_occ, decr_frac, force_fg)");
    CK(cudaSetDevice(h->device));
    request_refresh(h, frac, force_fg != 0);
    LVB_CATCH
}
int lvb_pawcs_refresh_model(lvb_handle h, uint32_t base_occ, float decr_frac, int force_fg) {
    LVB_TRY
    REQUIRE(h != nullptr, "null handle");
    REQUIRE(h->algo == LVB_ALGO_PAWCS, "not a PAWCS instance");
    CK(cudaSetDevice(h->device));
    paw_request_refresh(h, base_occ, decr_frac, force_fg != 0);
    LVB_CATCH
}
int lvb_set_auto_model_reset(lvb_handle h, int enabled) {
    LVB_TRY
    REQUIRE(h != nullptr, "null handle");
    REQUIRE(h->initialized, "algo must be initialized first");
    CK(cudaSetDevice(h->device));
    FrameCtl f; get_ctl(h, f); f.auto_reset = enabled ? 1 : 0; put_ctl(h, f);
    LVB_CATCH
}
int lvb_get_roi(lvb_handle h, uint8_t* out) {
    LVB_TRY
    REQUIRE(h && out, "null argument");
    REQUIRE(!h->roi_host.empty(), "no ROI set");
    std::memcpy(out, h->roi_host.data(), h->roi_host.size());
    LVB_CATCH
}
int lvb_set_roi(lvb_handle h, const uint8_t* roi) {
    LVB_TRY
    REQUIRE(h && roi, "provided ROI must be non-empty and of type 8UC1");
    REQUIRE(h->initialized, "setROI before initialize(): pass the ROI to lvb_initialize instead");
    // IIBackgroundSubtractor::setROI (BackgroundSubtractionUtils.cpp:38-48): validate, then re-initialise from the current background image
    const size_t npx = (size_t)h->W * h->H;
    size_t nz = 0;
    for(size_t i = 0; i < npx; ++i) nz += roi[i] != 0;
    REQUIRE(nz > 0, "provided ROI must have at least one valid pixel");
    std::vector<uint8_t> bg(npx * h->C), r(roi, roi + npx);
    get_bg_image(h, bg.data(), nullptr);
    do_initialize(h, bg.data(), h->W, h->H, h->C, (size_t)h->W * h->C, r.data());
    LVB_CATCH
}
int lvb_state_size(lvb_handle h, const char* name, size_t* bytes) {
    LVB_TRY
    REQUIRE(h && name && bytes, "null argument");
    REQUIRE(h->initialized, "algo must be initialized first");
    *bytes = state_bytes(h, name);
    LVB_CATCH
}
int lvb_state_get(lvb_handle h, const char* name, void* out, size_t bytes) {
    LVB_TRY
    REQUIRE(h && name && out, "null argument");
    state_get(h, name, out, bytes);
    LVB_CATCH
}
int lvb_state_set(lvb_handle h, const char* name, const void* in, size_t bytes) {
    LVB_TRY
    REQUIRE(h && name && in, "null argument");
    state_set(h, name, in, bytes);
    LVB_CATCH
}
int lvb_set_collect_stats(lvb_handle h, int enabled) {
    LVB_TRY
    REQUIRE(h != nullptr, "null handle");
    h->collect_stats = enabled ? 1 : 0;
    LVB_CATCH
}
int lvb_get_stats(lvb_handle h, uint64_t out[5]) {
    LVB_TRY
    REQUIRE(h && out, "null argument");
    REQUIRE(h->initialized, "algo must be initialized first");
    CK(cudaSetDevice(h->device));
    FrameCtl f; get_ctl(h, f);
    out[0] = (uint64_t)h->roi_count * h->stat_frames; out[1] = f.stat_scanned; out[2] = f.stat_writes; out[3] = f.stat_fg; out[4] = h->stat_frames;
    LVB_CATCH
}
uint64_t lvb_kernel_launch_count(void) { return g_launches.load(); }
int lvb_set_profile(lvb_handle h, int enabled) {
    LVB_TRY
    REQUIRE(h != nullptr, "null handle");
    h->profile = enabled != 0;
    LVB_CATCH
}
int lvb_get_profile(lvb_handle h, double* ms_total, uint64_t* launches) {
    LVB_TRY
    if(h && !h->trace.empty()) {
        CK(cudaSetDevice(h->device)); sync_streams(h);
        // completion time of every marked kernel relative to the start of ITS frame (marks sit on three streams), averaged over frames
        std::vector<std::pair<std::string, std::pair<double, int>>> agg;
        size_t start = 0;
        for(size_t k = 0; k < h->trace.size(); ++k) {
            if(std::string(h->trace[k].first) == "frame_start") { start = k; continue; }
            float ms = 0; cudaEventElapsedTime(&ms, h->trace[start].second, h->trace[k].second);
            bool found = false;
            for(auto& a : agg) if(a.first == h->trace[k].first) { a.second.first += ms; a.second.second++; found = true; }
            if(!found) agg.push_back({h->trace[k].first, {ms, 1}});
        }
        for(auto& a : agg) fprintf(stderr, "[lvb trace] done at +%8.1f us  %s\n", a.second.first / a.second.second * 1e3, a.first.c_str());
        for(auto& t : h->trace) cudaEventDestroy(t.second);
        h->trace.clear();
    }
    REQUIRE(h && ms_total && launches, "null argument");
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));
    for(size_t i = 0; i + 1 < h->prof_events.size(); i += 2) {
        float ms = 0; CK(cudaEventElapsedTime(&ms, h->prof_events[i], h->prof_events[i + 1]));
        h->prof_ms += ms; ++h->prof_n;
        cudaEventDestroy(h->prof_events[i]); cudaEventDestroy(h->prof_events[i + 1]);
    }
    h->prof_events.clear();
    *ms_total = h->prof_ms; *launches = h->prof_n;
    h->prof_ms = 0; h->prof_n = 0;
    LVB_CATCH
}
int lvb_get_profile_feedback(lvb_handle h, double* ms_total, uint64_t* launches) {
    LVB_TRY
    REQUIRE(h && ms_total && launches, "null argument");
    CK(cudaSetDevice(h->device));
    sync_streams(h);
    for(size_t i = 0; i + 1 < h->prof2_events.size(); i += 2) {
        float ms = 0; CK(cudaEventElapsedTime(&ms, h->prof2_events[i], h->prof2_events[i + 1]));
        h->prof2_ms += ms; ++h->prof2_n;
        cudaEventDestroy(h->prof2_events[i]); cudaEventDestroy(h->prof2_events[i + 1]);
    }
    h->prof2_events.clear();
    *ms_total = h->prof2_ms; *launches = h->prof2_n;
    h->prof2_ms = 0; h->prof2_n = 0;
    LVB_CATCH
}
int lvb_get_profile_tail(lvb_handle h, double* ms_total, uint64_t* launches) {
    LVB_TRY
    REQUIRE(h && ms_total && launches, "null argument");
    CK(cudaSetDevice(h->device));
    sync_streams(h);
    for(size_t i = 0; i + 1 < h->prof3_events.size(); i += 2) {
        float ms = 0; CK(cudaEventElapsedTime(&ms, h->prof3_events[i], h->prof3_events[i + 1]));
        h->prof3_ms += ms; ++h->prof3_n;
        cudaEventDestroy(h->prof3_events[i]); cudaEventDestroy(h->prof3_events[i + 1]);
    }
    h->prof3_events.clear();
    *ms_total = h->prof3_ms; *launches = h->prof3_n;
    h->prof3_ms = 0; h->prof3_n = 0;
    LVB_CATCH
}
int lvb_host_alloc(void** out, size_t bytes) {
    LVB_TRY
    REQUIRE(out != nullptr, "null argument");
    CK(cudaMallocHost(out, bytes));
    LVB_CATCH
}
int lvb_host_free(void* p) { if(p) cudaFreeHost(p); return 0; }


/// standalone mask operators on byte masks (non-zero = set): the bit-packed kernels the subtractors use, exposed so they
/// can be checked one by one against the OpenCV-equivalent oracle ops. op: 0 dilate(r) 1 erode(r) 2 median(k) 3 holes
/// (pixels NOT reached by cv::floodFill from (0,0), i.e. floodFill+bitwise_not) ; r in {1,3}
int lvb_mask_op(int op, const uint8_t* src, uint8_t* dst, int W, int H, int param, int device) {
    LVB_TRY
    REQUIRE(src && dst && W >= 1 && H >= 1 && W <= 8192, "bad mask arguments");
    REQUIRE(lvb_device_count() > 0, "no CUDA device available: litiv_b200 has no CPU fallback");
    CK(cudaSetDevice(device));
    const int WW = (W + 31) / 32, Wp = WW * 32;
    const size_t bp = (size_t)H * WW;
    std::vector<uint32_t> hb(bp, 0);
    for(int y = 0; y < H; ++y) for(int x = 0; x < W; ++x) if(src[(size_t)y * W + x]) hb[(size_t)y * WW + (x >> 5)] |= 1u << (x & 31);
    cudaStream_t st = 0;
    uint32_t* d = dalloc<uint32_t>(st, bp * 3);
    uint32_t* par = nullptr; ushort* rk = nullptr; uint8_t* dm = nullptr;
    try {
        h2d(st, d, hb.data(), bp * 4);
        uint32_t* out = d + bp;
        PostArgs P{};
        P.W = W; P.H = H; P.WW = WW; P.Wp = Wp;
        const dim3 wg((WW + 255) / 256, H);
        if(op == 0 || op == 1) {
            REQUIRE(param == 1 || param == 3, "morphology radius must be 1 or 3");
            mask_morph_kernel<<<wg, 256, 0, st>>>(d, out, W, H, WW, op == 0, param); LAUNCHED();
        } else if(op == 2) {
            REQUIRE(param >= 1 && param <= 31 && (param & 1), "median blur kernel size must be odd and <= 31");
            pp_median<<<dim3(Wp / 32, (H + 8 * MEDIAN_ROWS - 1) / (8 * MEDIAN_ROWS)), dim3(32, 8), 0, st>>>(d, out, nullptr, 0, W, H, WW, param); LAUNCHED();
        } else if(op == 3) {
            HoleArgs Hh{};
            Hh.W = W; Hh.H = H; Hh.WW = WW; Hh.RS = (W + 1) / 2 + 1; Hh.pre = d; Hh.raw = d + 2 * bp; Hh.comb = out;
            par = dalloc<uint32_t>(st, (size_t)H * Hh.RS + 1); rk = dalloc<ushort>(st, bp);
            Hh.parent = par; Hh.rankbase = rk;
            const int rb = (H + 7) / 8;
            pp_holes_init<<<rb, 256, 0, st>>>(Hh); LAUNCHED();
            pp_holes_union<<<rb, 256, 0, st>>>(Hh); LAUNCHED();
            pp_holes_only<<<rb, 256, 0, st>>>(Hh); LAUNCHED();
        } else throw std::runtime_error("unknown mask op");
        d2h(st, hb.data(), out, bp * 4);
    } catch(...) { cudaFree(d); cudaFree(par); cudaFree(rk); cudaFree(dm); throw; }
    cudaFree(d); cudaFree(par); cudaFree(rk); cudaFree(dm);
    for(int y = 0; y < H; ++y) for(int x = 0; x < W; ++x) dst[(size_t)y * W + x] = ((hb[(size_t)y * WW + (x >> 5)] >> (x & 31)) & 1u) ? 255 : 0;
    LVB_CATCH
}

int lvb_lbsp_compute(const uint8_t* img, const uint8_t* ref, int W, int H, int C, int use_rel, float rel, int thr, uint16_t* out, int device) {
    LVB_TRY
    REQUIRE(img && out && (C == 1 || C == 3), "input image must be non-empty, continuous, and of type 8UC1/8UC3");
    REQUIRE(W >= 5 && H >= 5, "input image size is too small to compute descriptors with current patch size");
    REQUIRE(!use_rel || rel >= 0, "lbsp internal relative threshold must be non-negative");
    REQUIRE(lvb_device_count() > 0, "no CUDA device available: litiv_b200 has no CPU fallback");
    CK(cudaSetDevice(device));
    const size_t pitch = ((size_t)W * C + 127) / 128 * 128, nout = (size_t)W * H * C;
    uint8_t* d_img = dalloc<uint8_t>((cudaStream_t)0, pitch * H), *d_ref = nullptr;
    uint16_t* d_out = nullptr;
    cudaError_t e = cudaSuccess;
    try {
        d_out = dalloc<uint16_t>((cudaStream_t)0, nout, false);
        CK(cudaMemcpy2D(d_img, pitch, img, (size_t)W * C, (size_t)W * C, H, cudaMemcpyHostToDevice));
        if(ref) { d_ref = dalloc<uint8_t>((cudaStream_t)0, pitch * H); CK(cudaMemcpy2D(d_ref, pitch, ref, (size_t)W * C, (size_t)W * C, H, cudaMemcpyHostToDevice)); }
        h2d((cudaStream_t)0, d_out, out, nout * 2); // the 2-px border keeps the caller's content, as in the reference
        CUtensorMap tmap;
        LbspArgs A{};
        A.W = W; A.H = H; A.img = d_img; A.ipitch = pitch; A.ref = d_ref; A.rpitch = pitch; A.out = d_out; A.use_rel = use_rel; A.rel = rel; A.thr = thr;
        A.use_tma = make_image_tmap(&tmap, d_img, W, H, C, pitch) ? 1 : 0;
        const dim3 g((W + TILE_W - 1) / TILE_W, (H + TILE_H - 1) / TILE_H), b(TILE_W, TILE_H);
        if(C == 1) lbsp_dense_kernel<1><<<g, b>>>(A, tmap); else lbsp_dense_kernel<3><<<g, b>>>(A, tmap);
        LAUNCHED();
        d2h((cudaStream_t)0, out, d_out, nout * 2);
    } catch(...) { cudaFree(d_img); cudaFree(d_ref); cudaFree(d_out); throw; }
    cudaFree(d_img); cudaFree(d_ref); cudaFree(d_out);
    (void)e;
    LVB_CATCH
}

int lvb_lbsp_gradient(const uint8_t* img, int W, int H, int C, uint8_t* out, int device) {
    LVB_TRY
    REQUIRE(img && out && C >= 1 && C <= 4, "input image must be non-empty, continuous, and of type 8UC1 .. 8UC4");   // LBSP.hpp:235: any channel count
    REQUIRE(W >= 5 && H >= 5, "input image size is too small to compute descriptors with current patch size");
    REQUIRE(lvb_device_count() > 0, "no CUDA device available: litiv_b200 has no CPU fallback");
    CK(cudaSetDevice(device));
    const size_t pitch = ((size_t)W * C + 127) / 128 * 128, npx = (size_t)W * H;
    uint8_t* d_img = dalloc<uint8_t>((cudaStream_t)0, pitch * H);
    uchar4* d_out = nullptr;
    try {
        d_out = dalloc<uchar4>((cudaStream_t)0, npx, false);
        CK(cudaMemcpy2D(d_img, pitch, img, (size_t)W * C, (size_t)W * C, H, cudaMemcpyHostToDevice));
        CUtensorMap tmap;
        LbspGradArgs A{};
        A.W = W; A.H = H; A.img = d_img; A.ipitch = pitch; A.out = d_out;
        A.use_tma = make_image_tmap(&tmap, d_img, W, H, C, pitch) ? 1 : 0;
        const dim3 g((W + TILE_W - 1) / TILE_W, (H + TILE_H - 1) / TILE_H), b(TILE_W, TILE_H);
        if(C == 1) lbsp_gradient_kernel<1><<<g, b>>>(A, tmap); else if(C == 2) lbsp_gradient_kernel<2><<<g, b>>>(A, tmap);
        else if(C == 3) lbsp_gradient_kernel<3><<<g, b>>>(A, tmap); else lbsp_gradient_kernel<4><<<g, b>>>(A, tmap);
        LAUNCHED();
        d2h((cudaStream_t)0, out, d_out, npx * 4);
    } catch(...) { cudaFree(d_img); cudaFree(d_out); throw; }
    cudaFree(d_img); cudaFree(d_out);
    LVB_CATCH
}

/// lv::BinClassif::accumulate (datasets/src/metrics.cpp:21-61) on the device; see csrc/metrics.cuh
static void binclassif_run(cudaStream_t st, int W, int H, int WW, const uint8_t* d_classif, const uint32_t* d_bits, const uint8_t* gt, const uint8_t* roi,
                           uint64_t counters[6], uint8_t* d_gt, uint8_t* d_roi, unsigned long long* d_cnt) {
    const size_t npx = (size_t)W * H;
    CK(cudaMemsetAsync(d_cnt, 0, BC_COUNT * sizeof(unsigned long long), st));
    CK(cudaMemcpyAsync(d_gt, gt, npx, cudaMemcpyHostToDevice, st));
    if(roi) CK(cudaMemcpyAsync(d_roi, roi, npx, cudaMemcpyHostToDevice, st));
    BinClassifArgs A{};
    A.W = W; A.H = H; A.WW = WW; A.classif = d_classif; A.cpitch = (size_t)W; A.classif_bits = d_bits; A.gt = d_gt; A.gpitch = (size_t)W;
    A.roi = roi ? d_roi : nullptr; A.rpitch = (size_t)W; A.counters = d_cnt;
    binclassif_kernel<<<dim3((W + 31) / 32, (H + 7) / 8), dim3(32, 8), 0, st>>>(A); LAUNCHED();
    unsigned long long h[BC_COUNT];
    d2h(st, h, d_cnt, sizeof(h));
    for(int i = 0; i < BC_COUNT; ++i) counters[i] += h[i];
}
int lvb_binclassif_accumulate(lvb_handle h, const uint8_t* gt, const uint8_t* roi, uint64_t counters[6]) {
    LVB_TRY
    REQUIRE(h != nullptr && counters != nullptr, "null argument");
    REQUIRE(h->initialized, "algo & model must be initialized first");
    CK(cudaSetDevice(h->device));
    const size_t npx = (size_t)h->W * h->H;
    if(!gt) { counters[BC_DC] += npx; return 0; } // metrics.cpp:26-29: no groundtruth -> every pixel is a don't-care
    while(sync_next(h)) {}
    if(h->s_post) CK(cudaStreamSynchronize(h->s_post));   // the final mask (lastfg) is written on the mask stream
    if(!h->eval_gt) {   // scratch kept with the instance: scoring runs once per frame
        h->eval_gt = dalloc<uint8_t>(h->stream, npx, false); h->eval_roi = dalloc<uint8_t>(h->stream, npx, false);
        h->eval_cnt = dalloc<unsigned long long>(h->stream, BC_COUNT);
    }
    // SuBSENSE: on the (now idle) mask stream, beside the feedback kernel still running on the instance stream
    binclassif_run(h->algo == LVB_ALGO_SUBSENSE && h->s_post ? h->s_post : h->stream, h->W, h->H, h->WW, nullptr, h->lastfg, gt, roi, counters,
                   h->eval_gt, h->eval_roi, h->eval_cnt);
    LVB_CATCH
}
int lvb_binclassif(const uint8_t* classif, const uint8_t* gt, const uint8_t* roi, int W, int H, uint64_t counters[6], int device) {
    LVB_TRY
    REQUIRE(classif && counters && W >= 1 && H >= 1, "binary classifier results must be non-empty and of type 8UC1");
    REQUIRE(lvb_device_count() > 0, "no CUDA device available: litiv_b200 has no CPU fallback");
    const size_t npx = (size_t)W * H;
    if(!gt) { counters[BC_DC] += npx; return 0; }
    CK(cudaSetDevice(device));
    uint8_t* d_c = nullptr, *d_gt = nullptr, *d_roi = nullptr; unsigned long long* d_cnt = nullptr;
    try {
        d_c = dalloc<uint8_t>((cudaStream_t)0, npx, false); d_gt = dalloc<uint8_t>((cudaStream_t)0, npx, false);
        d_roi = dalloc<uint8_t>((cudaStream_t)0, npx, false); d_cnt = dalloc<unsigned long long>((cudaStream_t)0, BC_COUNT);
        CK(cudaMemcpy(d_c, classif, npx, cudaMemcpyHostToDevice));
        binclassif_run((cudaStream_t)0, W, H, (W + 31) / 32, d_c, nullptr, gt, roi, counters, d_gt, d_roi, d_cnt);
    } catch(...) { cudaFree(d_c); cudaFree(d_gt); cudaFree(d_roi); cudaFree(d_cnt); throw; }
    cudaFree(d_c); cudaFree(d_gt); cudaFree(d_roi); cudaFree(d_cnt);
    LVB_CATCH
}
/// BinClassifMetrics (datasets/include/litiv/datasets/metrics.hpp:213-257): recall, specificity, FPR, FNR, PBC, precision, F-measure, MCC.
/// Host arithmetic on six integers; no device involved.
int lvb_binclassif_metrics(const uint64_t c[6], double out[8]) {
    LVB_TRY
    REQUIRE(c && out, "null argument");
    const double TP = (double)c[BC_TP], TN = (double)c[BC_TN], FP = (double)c[BC_FP], FN = (double)c[BC_FN];
    const uint64_t total = c[BC_TP] + c[BC_TN] + c[BC_FP] + c[BC_FN];
    const double recall = (c[BC_TP] + c[BC_FN]) > 0 ? TP / (double)(c[BC_TP] + c[BC_FN]) : 0;
    const double precision = (c[BC_TP] + c[BC_FP]) > 0 ? TP / (double)(c[BC_TP] + c[BC_FP]) : 0;
    out[0] = recall;
    out[1] = (c[BC_TN] + c[BC_FP]) > 0 ? TN / (double)(c[BC_TN] + c[BC_FP]) : 0;
    out[2] = (c[BC_FP] + c[BC_TN]) > 0 ? FP / (double)(c[BC_FP] + c[BC_TN]) : 0;
    out[3] = (c[BC_TP] + c[BC_FN]) > 0 ? FN / (double)(c[BC_TP] + c[BC_FN]) : 0;
    out[4] = total > 0 ? 100.0 * (double)(c[BC_FN] + c[BC_FP]) / (double)total : 0;
    out[5] = precision;
    out[6] = (recall + precision) > 0 ? 2.0 * (recall * precision) / (recall + precision) : 0;
    const bool ok = (c[BC_TP] + c[BC_FP]) > 0 && (c[BC_TP] + c[BC_FN]) > 0 && (c[BC_TN] + c[BC_FP]) > 0 && (c[BC_TN] + c[BC_FN]) > 0;
    out[7] = ok ? ((TP * TN) - (double)(c[BC_FP] * c[BC_FN])) / std::sqrt((TP + FP) * (double)(c[BC_TP] + c[BC_FN]) * (double)(c[BC_TN] + c[BC_FP]) * (double)(c[BC_TN] + c[BC_FN])) : 0;
    LVB_CATCH
}

} // extern "C"

#include "vibe_host.cuh"
#include "pbas_host.cuh"
#include "edge_host.cuh"
