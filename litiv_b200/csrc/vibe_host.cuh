// litiv_b200 — host side of the ViBe entry points (include/litiv_b200.h: lvb_vibe_*). Included at the end of litiv_b200.cu
// (same translation unit: CK / REQUIRE / LVB_TRY, dalloc, is_pinned, lr_to_fixed come from there).
// The reference's ViBe classes derive from cv::BackgroundSubtractor directly (video/include/litiv/video/BackgroundSubtractorViBe.hpp:
// 50-77): no ROI, no LBSP layer, initialize(img) only, so they get their own small context instead of a fourth lvb_context flavour.
#pragma once
#include "vibe.cuh"

struct lvb_vibe_context {
    int device = 0, MC = 3, thr = 20, N = 20, REQ = 2;
    uint64_t seed = 0;
    int W = 0, H = 0, Wp = 0;
    size_t plane = 0, ipitch = 0;
    bool initialized = false;
    uint32_t frame = 0;                 // frames applied since initialize (Philox counter word 0; the first apply is frame 1)
    cudaStream_t stream = nullptr;
    uint8_t *d_img = nullptr, *d_mask = nullptr, *h_img = nullptr, *h_mask = nullptr;
    void* bg = nullptr;
    // neighbour writes queued by frame k are applied by the scan of frame k+1: planes [k & 1] are written, [(k & 1) ^ 1] read
    ushort* intents[2] = {nullptr, nullptr}; void* nbcol[2] = {nullptr, nullptr};
    bool nb_pending = false;           // the planes of the latest frame still hold unapplied writes
    unsigned long long* d_stats = nullptr; int collect_stats = 0; uint64_t stat_frames = 0;
    bool profile = false; std::vector<cudaEvent_t> prof_events; double prof_ms = 0; uint64_t prof_n = 0;

    void free_all() {
        for(void* p : {(void*)d_img, (void*)d_mask, bg, (void*)intents[0], (void*)intents[1], nbcol[0], nbcol[1], (void*)d_stats}) if(p) cudaFree(p);
        if(h_img) cudaFreeHost(h_img);
        if(h_mask) cudaFreeHost(h_mask);
        d_img = d_mask = h_img = h_mask = nullptr; bg = nullptr; intents[0] = intents[1] = nullptr; nbcol[0] = nbcol[1] = nullptr; d_stats = nullptr;
        nb_pending = false;
        initialized = false;
    }
};

namespace {

VibeArgs vibe_args(lvb_vibe_context* c, const uint8_t* d_img, size_t pitch, int in_ch, uint8_t* d_mask, double lr) {
    VibeArgs A{};
    A.W = c->W; A.H = c->H; A.Wp = c->Wp; A.N = c->N; A.REQ = c->REQ;
    A.thr = (uint32_t)(c->MC == 1 ? c->thr : (c->thr * 3) * (c->thr * 3)); // threshold <= 255: (3*thr)^2 < 2^20
    A.img = d_img; A.ipitch = pitch; A.in_ch = in_ch;
    A.bg = c->bg; A.plane = c->plane; A.mask = d_mask; A.mpitch = (size_t)c->W;
    const int cur = (int)(c->frame & 1u); // planes written by frame c->frame
    A.intents = c->intents[cur]; A.nbcol = c->nbcol[cur];
    A.prev_intents = c->nb_pending ? c->intents[cur ^ 1] : nullptr; A.prev_nbcol = c->nbcol[cur ^ 1];
    A.frame = c->frame; A.seed = c->seed; A.lr = lr_to_fixed(lr);
    A.lr_magic = magic_of(A.lr); A.n_magic = magic_of((uint32_t)c->N);
    A.stats = c->collect_stats ? c->d_stats : nullptr;
    return A;
}
dim3 vibe_grid(const lvb_vibe_context* c) { return dim3((c->W + 31) / 32, (c->H + 7) / 8); }

void vibe_check_image(const lvb_vibe_context* c, const void* img, int channels) {
    REQUIRE(img != nullptr, "input image must be non-empty");
    if(c->MC == 1) REQUIRE(channels == 1, "input image type must be 8UC1 and match the initialization size");           // ViBe.cpp:61, :83
    else REQUIRE(channels == 1 || channels == 3, "input image type must be 8UC1 or 8UC3 and match the initialization size"); // :118, :144
}

void vibe_enqueue(lvb_vibe_context* c, const uint8_t* d_img, size_t pitch, int in_ch, uint8_t* d_mask, double lr) {
    c->frame += 1;
    const VibeArgs A = vibe_args(c, d_img, pitch, in_ch, d_mask, lr);
    const dim3 g = vibe_grid(c), b(32, 8);
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if(c->profile) { CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventRecord(e0, c->stream)); }
    if(c->MC == 1) vibe_phaseA<1><<<g, b, 0, c->stream>>>(A); else vibe_phaseA<3><<<g, b, 0, c->stream>>>(A);
    LAUNCHED();
    if(c->profile) { CK(cudaEventRecord(e1, c->stream)); c->prof_events.push_back(e0); c->prof_events.push_back(e1); }
    c->nb_pending = true; // this frame's neighbour writes wait for the next frame's scan (or vibe_flush_pending)
    if(c->collect_stats) ++c->stat_frames;
}

/// apply the neighbour writes the latest frame queued (model export, getBackgroundImage, model import)
void vibe_flush_pending(lvb_vibe_context* c) {
    if(!c->nb_pending) return;
    VibeArgs A = vibe_args(c, c->d_img, c->ipitch, c->MC, c->d_mask, 1.0);
    const int last = (int)(c->frame & 1u); // planes the latest frame wrote
    A.prev_intents = c->intents[last]; A.prev_nbcol = c->nbcol[last];
    if(c->MC == 1) vibe_phaseB<1><<<vibe_grid(c), dim3(32, 8), 0, c->stream>>>(A); else vibe_phaseB<3><<<vibe_grid(c), dim3(32, 8), 0, c->stream>>>(A);
    LAUNCHED();
    c->nb_pending = false;
}

} // namespace

extern "C" {

int lvb_vibe_create(int model_channels, int color_dist_threshold, int n_samples, int n_required, int device, uint64_t seed, lvb_vibe_handle* out) {
    LVB_TRY
    REQUIRE(out != nullptr, "null output handle");
    REQUIRE(model_channels == 1 || model_channels == 3, "ViBe model must have 1 or 3 channels");
    REQUIRE(n_samples > 0 && n_required <= n_samples, "algo cannot require more sample matches than sample count in model"); // ViBe.cpp:27
    REQUIRE(n_samples <= 255, "at most 255 samples per pixel are supported");
    REQUIRE(color_dist_threshold >= 0 && color_dist_threshold <= 255, "colour distance threshold must be in [0,255]");
    const int ndev = lvb_device_count();
    REQUIRE(ndev > 0, "no CUDA device available: litiv_b200 has no CPU fallback");
    REQUIRE(device >= 0 && device < ndev, "invalid CUDA device id");
    CK(cudaSetDevice(device));
    lvb_vibe_context* c = new lvb_vibe_context();
    c->device = device; c->MC = model_channels; c->thr = color_dist_threshold; c->N = n_samples; c->REQ = n_required; c->seed = seed;
    cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if(e != cudaSuccess) { delete c; CK(e); }
    *out = c;
    LVB_CATCH
}

int lvb_vibe_destroy(lvb_vibe_handle h) {
    if(!h) return 0;
    cudaSetDevice(h->device);
    if(h->stream) cudaStreamSynchronize(h->stream);
    for(cudaEvent_t e : h->prof_events) cudaEventDestroy(e);
    h->free_all();
    if(h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return 0;
}

int lvb_vibe_initialize(lvb_vibe_handle h, const uint8_t* img, int W, int H, int channels, size_t step) {
    LVB_TRY
    REQUIRE(h != nullptr, "null handle");
    REQUIRE(img != nullptr && W > 0 && H > 0, "provided image for initialization must be non-empty and continuous"); // ViBe.cpp:59-60
    vibe_check_image(h, img, channels);
    REQUIRE(step >= (size_t)W * channels, "row step smaller than a row");
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));
    h->free_all();
    h->W = W; h->H = H; h->Wp = (W + 31) / 32 * 32; h->plane = (size_t)H * h->Wp;
    h->ipitch = ((size_t)W * h->MC + 127) / 128 * 128;
    h->d_img = dalloc<uint8_t>(h->stream, h->ipitch * H);
    h->d_mask = dalloc<uint8_t>(h->stream, (size_t)W * H);
    CK(cudaMallocHost((void**)&h->h_img, (size_t)W * H * h->MC));
    CK(cudaMallocHost((void**)&h->h_mask, (size_t)W * H));
    h->bg = dalloc<uint8_t>(h->stream, (size_t)h->N * h->plane * (h->MC == 1 ? 1 : 4));
    for(int i = 0; i < 2; ++i) { h->intents[i] = dalloc<ushort>(h->stream, h->plane); h->nbcol[i] = dalloc<uint8_t>(h->stream, h->plane * (h->MC == 1 ? 1 : 4)); }
    h->d_stats = dalloc<unsigned long long>(h->stream, 3);
    h->frame = 0; h->stat_frames = 0;
    CK(cudaMemcpy2DAsync(h->d_img, h->ipitch, img, step, (size_t)W * channels, H, cudaMemcpyHostToDevice, h->stream));
    const VibeArgs A = vibe_args(h, h->d_img, h->ipitch, channels, h->d_mask, 1.0);
    if(h->MC == 1) vibe_init_kernel<1><<<vibe_grid(h), dim3(32, 8), 0, h->stream>>>(A); else vibe_init_kernel<3><<<vibe_grid(h), dim3(32, 8), 0, h->stream>>>(A);
    LAUNCHED();
    CK(cudaStreamSynchronize(h->stream));
    h->initialized = true;
    LVB_CATCH
}

int lvb_vibe_apply(lvb_vibe_handle h, const uint8_t* img, int channels, uint8_t* fgmask, double lr) {
    LVB_TRY
    REQUIRE(h != nullptr && h->initialized, "algo must be initialized first"); // ViBe.cpp:79
    REQUIRE(lr > 0, "learning rate must be a positive value");               // :80
    vibe_check_image(h, img, channels);
    REQUIRE(fgmask != nullptr, "output mask must be provided");
    CK(cudaSetDevice(h->device));
    const size_t row = (size_t)h->W * channels, npx = (size_t)h->W * h->H;
    const uint8_t* src = img;
    if(!is_pinned(img)) { std::memcpy(h->h_img, img, row * h->H); src = h->h_img; }
    CK(cudaMemcpy2DAsync(h->d_img, h->ipitch, src, row, row, h->H, cudaMemcpyHostToDevice, h->stream));
    vibe_enqueue(h, h->d_img, h->ipitch, channels, h->d_mask, lr);
    const bool direct = is_pinned(fgmask);
    CK(cudaMemcpyAsync(direct ? fgmask : h->h_mask, h->d_mask, npx, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if(!direct) std::memcpy(fgmask, h->h_mask, npx);
    LVB_CATCH
}

int lvb_vibe_apply_device(lvb_vibe_handle h, const uint8_t* d_img, int channels, size_t d_step, uint8_t* d_fgmask, double lr) {
    LVB_TRY
    REQUIRE(h != nullptr && h->initialized, "algo must be initialized first");
    REQUIRE(lr > 0, "learning rate must be a positive value");
    vibe_check_image(h, d_img, channels);
    REQUIRE(d_step >= (size_t)h->W * channels, "row step smaller than a row");
    CK(cudaSetDevice(h->device));
    vibe_enqueue(h, d_img, d_step, channels, d_fgmask ? d_fgmask : h->d_mask, lr);
    LVB_CATCH
}

int lvb_vibe_sync(lvb_vibe_handle h) {
    LVB_TRY
    REQUIRE(h != nullptr, "null handle");
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));
    LVB_CATCH
}

int lvb_vibe_get_background_image(lvb_vibe_handle h, uint8_t* out) {
    LVB_TRY
    REQUIRE(h != nullptr && h->initialized, "algo must be initialized first"); // ViBe.cpp:33
    REQUIRE(out != nullptr, "null output");
    CK(cudaSetDevice(h->device));
    const size_t n = (size_t)h->W * h->H * h->MC;
    vibe_flush_pending(h);
    uint8_t* d = dalloc<uint8_t>(h->stream, n, false);
    const VibeArgs A = vibe_args(h, h->d_img, h->ipitch, h->MC, h->d_mask, 1.0);
    if(h->MC == 1) vibe_background_kernel<1><<<vibe_grid(h), dim3(32, 8), 0, h->stream>>>(A, d); else vibe_background_kernel<3><<<vibe_grid(h), dim3(32, 8), 0, h->stream>>>(A, d);
    ++g_launches;
    cudaError_t e = cudaGetLastError();
    if(e == cudaSuccess) e = cudaMemcpyAsync(out, d, n, cudaMemcpyDeviceToHost, h->stream);
    if(e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(d);
    CK(e);
    LVB_CATCH
}

/* model samples in the reference's layout [N][H][W][C] (m_voBGImg, ViBe.hpp:67); set != 0 imports, and `frame` then replaces the
 * frame counter that indexes the Philox stream */
int lvb_vibe_model(lvb_vibe_handle h, uint8_t* inout, size_t bytes, int set, uint32_t frame) {
    LVB_TRY
    REQUIRE(h != nullptr && h->initialized, "algo must be initialized first");
    const size_t n = (size_t)h->N * h->W * h->H * h->MC;
    REQUIRE(inout != nullptr && bytes == n, "size mismatch for the ViBe model");
    CK(cudaSetDevice(h->device));
    if(set) h->nb_pending = false; else vibe_flush_pending(h); // an imported model is a complete one: queued writes are dropped
    uint8_t* d = dalloc<uint8_t>(h->stream, n, false);
    cudaError_t e = cudaSuccess;
    if(set) e = cudaMemcpyAsync(d, inout, n, cudaMemcpyHostToDevice, h->stream);
    const VibeArgs A = vibe_args(h, h->d_img, h->ipitch, h->MC, h->d_mask, 1.0);
    const dim3 g((h->W + 31) / 32, (h->H + 7) / 8, h->N);
    if(h->MC == 1) vibe_model_copy_kernel<1><<<g, dim3(32, 8), 0, h->stream>>>(A, d, set); else vibe_model_copy_kernel<3><<<g, dim3(32, 8), 0, h->stream>>>(A, d, set);
    ++g_launches;
    if(e == cudaSuccess) e = cudaGetLastError();
    if(e == cudaSuccess && !set) e = cudaMemcpyAsync(inout, d, n, cudaMemcpyDeviceToHost, h->stream);
    if(e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(d);
    CK(e);
    if(set) h->frame = frame;
    LVB_CATCH
}

int lvb_vibe_set_collect_stats(lvb_vibe_handle h, int enabled) {
    LVB_TRY
    REQUIRE(h != nullptr, "null handle");
    h->collect_stats = enabled != 0;
    LVB_CATCH
}
/* out[0..4] = px, samples_scanned, sample_writes, fg_px, frames (accumulated while enabled, since initialize) */
int lvb_vibe_get_stats(lvb_vibe_handle h, uint64_t out[5]) {
    LVB_TRY
    REQUIRE(h != nullptr && h->initialized && out, "algo must be initialized first");
    CK(cudaSetDevice(h->device));
    unsigned long long s[3];
    CK(cudaMemcpyAsync(s, h->d_stats, sizeof(s), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    out[0] = (uint64_t)h->W * h->H * h->stat_frames; out[1] = s[0]; out[2] = s[1]; out[3] = s[2]; out[4] = h->stat_frames;
    LVB_CATCH
}
int lvb_vibe_set_profile(lvb_vibe_handle h, int enabled) {
    LVB_TRY
    REQUIRE(h != nullptr, "null handle");
    h->profile = enabled != 0;
    LVB_CATCH
}
/* per-launch CUDA-event time of the dominant kernel (vibe_phaseA) on the instance's stream: read + reset */
int lvb_vibe_get_profile(lvb_vibe_handle h, double* ms_total, uint64_t* launches) {
    LVB_TRY
    REQUIRE(h != nullptr && ms_total && launches, "null argument");
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
void* lvb_vibe_stream(lvb_vibe_handle h) { return h ? (void*)h->stream : nullptr; }

} // extern "C"
