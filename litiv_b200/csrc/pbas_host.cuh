// litiv_b200 — host side of the PBAS entry points (include/litiv_b200.h: lvb_pbas_*). Included at the end of litiv_b200.cu, after
// vibe_host.cuh (same translation unit). Like ViBe, the reference's PBAS classes derive from cv::BackgroundSubtractor directly
// (video/include/litiv/video/BackgroundSubtractorPBAS.hpp:88-121): no ROI, initialize(img) only.
#pragma once
#include "pbas.cuh"

struct lvb_pbas_context {
    int device = 0, MC = 3, thr = 30, N = 35, REQ = 2;
    float t0 = 16.0f;
    uint64_t seed = 0;
    int W = 0, H = 0, Wp = 0, WW = 0;
    size_t plane = 0, ipitch = 0;
    bool initialized = false;
    uint32_t frame = 0;
    cudaStream_t stream = nullptr;
    uint8_t *d_img = nullptr, *d_mask = nullptr, *h_img = nullptr, *h_mask = nullptr;
    void* bg = nullptr;
    // planes written by frame k: [k & 1]; the scan of frame k+1 applies the neighbour writes they describe
    void *grad[2] = {nullptr, nullptr}, *col[2] = {nullptr, nullptr}; ushort* intents[2] = {nullptr, nullptr};
    bool nb_pending = false;
    float *R = nullptr, *T = nullptr, *meanmin = nullptr;
    uint32_t *raw_bits = nullptr, *fg_bits = nullptr, *magic = nullptr;
    PbasCtl* ctl = nullptr;
    unsigned long long* d_stats = nullptr; int collect_stats = 0; uint64_t stat_frames = 0;
    bool profile = false; std::vector<cudaEvent_t> prof_events; double prof_ms = 0; uint64_t prof_n = 0;

    size_t rec_bytes() const { return MC == 1 ? 2 : 8; }
    size_t grad_bytes() const { return MC == 1 ? 1 : 4; }
    void free_all() {
        for(void* p : {(void*)d_img, (void*)d_mask, bg, grad[0], grad[1], col[0], col[1], (void*)intents[0], (void*)intents[1], (void*)R, (void*)T, (void*)meanmin, (void*)raw_bits, (void*)fg_bits,
                       (void*)magic, (void*)ctl, (void*)d_stats}) if(p) cudaFree(p);
        if(h_img) cudaFreeHost(h_img);
        if(h_mask) cudaFreeHost(h_mask);
        d_img = d_mask = h_img = h_mask = nullptr; bg = nullptr; grad[0] = grad[1] = col[0] = col[1] = nullptr; intents[0] = intents[1] = nullptr;
        R = T = meanmin = nullptr; nb_pending = false;
        raw_bits = fg_bits = magic = nullptr; ctl = nullptr; d_stats = nullptr;
        initialized = false;
    }
};

namespace {

PbasArgs pbas_args(lvb_pbas_context* c, const uint8_t* d_img, size_t pitch, int in_ch, double lr) {
    PbasArgs A{};
    A.W = c->W; A.H = c->H; A.Wp = c->Wp; A.WW = c->WW; A.N = c->N; A.REQ = c->REQ; A.thr0 = (float)c->thr;
    A.img = d_img; A.ipitch = pitch; A.in_ch = in_ch;
    A.bg = c->bg; A.plane = c->plane; A.R = c->R; A.T = c->T; A.meanmin = c->meanmin;
    const int cur = (int)(c->frame & 1u);
    A.grad = c->grad[cur]; A.col = c->col[cur]; A.intents = c->intents[cur];
    A.prev_intents = c->nb_pending ? c->intents[cur ^ 1] : nullptr; A.prev_col = c->col[cur ^ 1]; A.prev_grad = c->grad[cur ^ 1];
    A.raw_bits = c->raw_bits; A.ctl = c->ctl;
    A.frame = c->frame; A.seed = c->seed; A.lr_override = lr_to_fixed(lr);
    A.n_magic = magic_of((uint32_t)c->N); A.magic = c->magic;
    A.stats = c->collect_stats ? c->d_stats : nullptr;
    return A;
}
dim3 pbas_grid(const lvb_pbas_context* c) { return dim3((c->W + 31) / 32, (c->H + 7) / 8); }

void pbas_check_image(const lvb_pbas_context* c, const void* img, int channels) {
    REQUIRE(img != nullptr, "input image must be non-empty");
    if(c->MC == 1) REQUIRE(channels == 1, "input image type must be 8UC1 and match the initialization size");           // PBAS.cpp:64, :116
    else REQUIRE(channels == 1 || channels == 3, "input image type must be 8UC1 or 8UC3 and match the initialization size"); // :288, :333
}

void pbas_enqueue(lvb_pbas_context* c, const uint8_t* d_img, size_t pitch, int in_ch, uint8_t* d_mask, double lr) {
    c->frame += 1;
    const PbasArgs A = pbas_args(c, d_img, pitch, in_ch, lr);
    const dim3 g = pbas_grid(c), b(32, 8);
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if(c->profile) { CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventRecord(e0, c->stream)); }
    if(c->MC == 1) pbas_phaseA<1><<<g, b, 0, c->stream>>>(A); else pbas_phaseA<3><<<g, b, 0, c->stream>>>(A);
    LAUNCHED();
    if(c->profile) { CK(cudaEventRecord(e1, c->stream)); c->prof_events.push_back(e0); c->prof_events.push_back(e1); }
    c->nb_pending = true; // this frame's self-diffusion writes wait for the next frame's scan (or pbas_flush_pending)
    const dim3 mg(c->Wp / 32, (c->H + 8 * MEDIAN_ROWS - 1) / (8 * MEDIAN_ROWS));
    pp_median<<<mg, b, 0, c->stream>>>(c->raw_bits, c->fg_bits, d_mask, (size_t)c->W, c->W, c->H, c->WW, 9); // PBAS.cpp:269 / :494
    LAUNCHED();
    if(c->collect_stats) ++c->stat_frames;
}

/// apply the neighbour writes the latest frame queued (state export, getBackgroundImage)
void pbas_flush_pending(lvb_pbas_context* c) {
    if(!c->nb_pending) return;
    PbasArgs A = pbas_args(c, c->d_img, c->ipitch, c->MC, 0.0);
    const int last = (int)(c->frame & 1u);
    A.prev_intents = c->intents[last]; A.prev_col = c->col[last]; A.prev_grad = c->grad[last];
    if(c->MC == 1) pbas_phaseB<1><<<pbas_grid(c), dim3(32, 8), 0, c->stream>>>(A); else pbas_phaseB<3><<<pbas_grid(c), dim3(32, 8), 0, c->stream>>>(A);
    LAUNCHED();
    c->nb_pending = false;
}

} // namespace

extern "C" {

int lvb_pbas_create(int model_channels, int init_color_dist_threshold, float init_update_rate, int n_samples, int n_required, int device,
                    uint64_t seed, lvb_pbas_handle* out) {
    LVB_TRY
    REQUIRE(out != nullptr, "null output handle");
    REQUIRE(model_channels == 1 || model_channels == 3, "PBAS model must have 1 or 3 channels");
    REQUIRE(n_samples > 0 && n_required <= n_samples, "algo cannot require more sample matches than sample count in model"); // PBAS.cpp:31
    REQUIRE(n_samples <= 255, "at most 255 samples per pixel are supported");
    REQUIRE(init_update_rate > 0 && init_update_rate <= 255, "default update rate must be in ]0,255]");                        // :32
    REQUIRE(init_color_dist_threshold >= 0, "colour distance threshold must be non-negative");
    const int ndev = lvb_device_count();
    REQUIRE(ndev > 0, "no CUDA device available: litiv_b200 has no CPU fallback");
    REQUIRE(device >= 0 && device < ndev, "invalid CUDA device id");
    CK(cudaSetDevice(device));
    lvb_pbas_context* c = new lvb_pbas_context();
    c->device = device; c->MC = model_channels; c->thr = init_color_dist_threshold; c->t0 = init_update_rate; c->N = n_samples; c->REQ = n_required;
    c->seed = seed;
    cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if(e != cudaSuccess) { delete c; CK(e); }
    *out = c;
    LVB_CATCH
}

int lvb_pbas_destroy(lvb_pbas_handle h) {
    if(!h) return 0;
    cudaSetDevice(h->device);
    if(h->stream) cudaStreamSynchronize(h->stream);
    for(cudaEvent_t e : h->prof_events) cudaEventDestroy(e);
    h->free_all();
    if(h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return 0;
}

int lvb_pbas_initialize(lvb_pbas_handle h, const uint8_t* img, int W, int H, int channels, size_t step) {
    LVB_TRY
    REQUIRE(h != nullptr, "null handle");
    REQUIRE(img != nullptr && W > 0 && H > 0, "provided image for initialization must be non-empty and continuous"); // PBAS.cpp:61-62
    pbas_check_image(h, img, channels);
    REQUIRE(step >= (size_t)W * channels, "row step smaller than a row");
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));
    h->free_all();
    h->W = W; h->H = H; h->Wp = (W + 31) / 32 * 32; h->WW = h->Wp / 32; h->plane = (size_t)H * h->Wp;
    h->ipitch = ((size_t)W * h->MC + 127) / 128 * 128;
    cudaStream_t st = h->stream;
    h->d_img = dalloc<uint8_t>(st, h->ipitch * H);
    h->d_mask = dalloc<uint8_t>(st, (size_t)W * H);
    CK(cudaMallocHost((void**)&h->h_img, (size_t)W * H * h->MC));
    CK(cudaMallocHost((void**)&h->h_mask, (size_t)W * H));
    h->bg = dalloc<uint8_t>(st, (size_t)h->N * h->plane * h->rec_bytes());
    for(int i = 0; i < 2; ++i) {
        h->grad[i] = dalloc<uint8_t>(st, h->plane * h->grad_bytes()); h->col[i] = dalloc<uint8_t>(st, h->plane * h->grad_bytes());
        h->intents[i] = dalloc<ushort>(st, h->plane);
    }
    h->R = dalloc<float>(st, h->plane); h->T = dalloc<float>(st, h->plane); h->meanmin = dalloc<float>(st, h->plane);
    h->raw_bits = dalloc<uint32_t>(st, (size_t)H * h->WW); h->fg_bits = dalloc<uint32_t>(st, (size_t)H * h->WW);
    h->d_stats = dalloc<unsigned long long>(st, 3);
    h->ctl = dalloc<PbasCtl>(st, 1);
    {
        uint32_t mg[257];
        mg[0] = 0; mg[1] = 0xFFFFFFFFu;
        for(uint32_t n = 2; n <= 256; ++n) mg[n] = (uint32_t)(0x100000000ull / n);
        h->magic = dalloc<uint32_t>(st, 257, false);
        h2d(st, h->magic, mg, sizeof(mg));
        PbasCtl c0{}; c0.former = 20.0f; // m_fFormerMeanGradDist(20), PBAS.cpp:29
        h2d(st, h->ctl, &c0, sizeof(c0));
    }
    h->frame = 0; h->stat_frames = 0;
    CK(cudaMemcpy2DAsync(h->d_img, h->ipitch, img, step, (size_t)W * channels, H, cudaMemcpyHostToDevice, st));
    const PbasArgs A = pbas_args(h, h->d_img, h->ipitch, channels, 0.0);
    const dim3 g = pbas_grid(h), b(32, 8);
    if(h->MC == 1) pbas_grad_kernel<1><<<g, b, 0, st>>>(A); else pbas_grad_kernel<3><<<g, b, 0, st>>>(A);
    LAUNCHED();
    if(h->MC == 1) pbas_init_kernel<1><<<g, b, 0, st>>>(A, h->t0); else pbas_init_kernel<3><<<g, b, 0, st>>>(A, h->t0);
    LAUNCHED();
    CK(cudaStreamSynchronize(st));
    h->initialized = true;
    LVB_CATCH
}

int lvb_pbas_apply(lvb_pbas_handle h, const uint8_t* img, int channels, uint8_t* fgmask, double lr) {
    LVB_TRY
    REQUIRE(h != nullptr && h->initialized, "algo must be initialized first"); // PBAS.cpp:113
    REQUIRE(!std::isnan(lr), "learning rate must not be NaN");
    pbas_check_image(h, img, channels);
    REQUIRE(fgmask != nullptr, "output mask must be provided");
    CK(cudaSetDevice(h->device));
    const size_t row = (size_t)h->W * channels, npx = (size_t)h->W * h->H;
    const uint8_t* src = img;
    if(!is_pinned(img)) { std::memcpy(h->h_img, img, row * h->H); src = h->h_img; }
    CK(cudaMemcpy2DAsync(h->d_img, h->ipitch, src, row, row, h->H, cudaMemcpyHostToDevice, h->stream));
    pbas_enqueue(h, h->d_img, h->ipitch, channels, h->d_mask, lr);
    const bool direct = is_pinned(fgmask);
    CK(cudaMemcpyAsync(direct ? fgmask : h->h_mask, h->d_mask, npx, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if(!direct) std::memcpy(fgmask, h->h_mask, npx);
    LVB_CATCH
}

int lvb_pbas_apply_device(lvb_pbas_handle h, const uint8_t* d_img, int channels, size_t d_step, uint8_t* d_fgmask, double lr) {
    LVB_TRY
    REQUIRE(h != nullptr && h->initialized, "algo must be initialized first");
    REQUIRE(!std::isnan(lr), "learning rate must not be NaN");
    pbas_check_image(h, d_img, channels);
    REQUIRE(d_step >= (size_t)h->W * channels, "row step smaller than a row");
    CK(cudaSetDevice(h->device));
    pbas_enqueue(h, d_img, d_step, channels, d_fgmask ? d_fgmask : h->d_mask, lr);
    LVB_CATCH
}

int lvb_pbas_sync(lvb_pbas_handle h) {
    LVB_TRY
    REQUIRE(h != nullptr, "null handle");
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));
    LVB_CATCH
}

int lvb_pbas_get_background_image(lvb_pbas_handle h, uint8_t* out) {
    LVB_TRY
    REQUIRE(h != nullptr && h->initialized, "algo must be initialized first"); // PBAS.cpp:38
    REQUIRE(out != nullptr, "null output");
    CK(cudaSetDevice(h->device));
    const size_t n = (size_t)h->W * h->H * h->MC;
    pbas_flush_pending(h);
    uint8_t* d = dalloc<uint8_t>(h->stream, n, false);
    const PbasArgs A = pbas_args(h, h->d_img, h->ipitch, h->MC, 0.0);
    if(h->MC == 1) pbas_background_kernel<1><<<pbas_grid(h), dim3(32, 8), 0, h->stream>>>(A, d); else pbas_background_kernel<3><<<pbas_grid(h), dim3(32, 8), 0, h->stream>>>(A, d);
    ++g_launches;
    cudaError_t e = cudaGetLastError();
    if(e == cudaSuccess) e = cudaMemcpyAsync(out, d, n, cudaMemcpyDeviceToHost, h->stream);
    if(e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(d);
    CK(e);
    LVB_CATCH
}

/* named state in the reference's layout: "bg_color" / "bg_grad" [N][H][W][C] u8 (m_voBGImg / m_voBGGrad), "R" / "T" / "meanmin" [H][W]
 * f32 (m_oDistThresholdFrame / m_oUpdateRateFrame / m_oMeanMinDistFrame), "rawmask" [H][W] u8 (before the median), "lastgrad" [H][W][C]
 * u8 (gradient magnitude of the latest frame), "scalars" 2 x f64 = frame counter (Philox index), m_fFormerMeanGradDist */
int lvb_pbas_state(lvb_pbas_handle h, const char* name, void* inout, size_t bytes, int set) {
    LVB_TRY
    REQUIRE(h != nullptr && h->initialized, "algo must be initialized first");
    REQUIRE(name && inout, "null argument");
    CK(cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    pbas_flush_pending(h); // the model buffers below are complete ones
    CK(cudaStreamSynchronize(st));
    const std::string n(name);
    const size_t npx = (size_t)h->W * h->H;
    if(n == "bg_color" || n == "bg_grad") {
        const size_t sz = (size_t)h->N * npx * h->MC;
        REQUIRE(bytes == sz, "size mismatch for state buffer " + n);
        uint8_t* d = dalloc<uint8_t>(st, sz, false);
        cudaError_t e = cudaSuccess;
        if(set) e = cudaMemcpyAsync(d, inout, sz, cudaMemcpyHostToDevice, st);
        const PbasArgs A = pbas_args(h, h->d_img, h->ipitch, h->MC, 0.0);
        const dim3 g((h->W + 31) / 32, (h->H + 7) / 8, h->N);
        const int which = n == "bg_grad";
        if(h->MC == 1) pbas_model_copy_kernel<1><<<g, dim3(32, 8), 0, st>>>(A, d, which, set); else pbas_model_copy_kernel<3><<<g, dim3(32, 8), 0, st>>>(A, d, which, set);
        ++g_launches;
        if(e == cudaSuccess) e = cudaGetLastError();
        if(e == cudaSuccess && !set) e = cudaMemcpyAsync(inout, d, sz, cudaMemcpyDeviceToHost, st);
        if(e == cudaSuccess) e = cudaStreamSynchronize(st);
        cudaFree(d);
        CK(e);
    } else if(n == "R" || n == "T" || n == "meanmin") {
        REQUIRE(bytes == npx * 4, "size mismatch for state buffer " + n);
        float* p = n == "R" ? h->R : n == "T" ? h->T : h->meanmin;
        if(set) CK(cudaMemcpy2DAsync(p, (size_t)h->Wp * 4, inout, (size_t)h->W * 4, (size_t)h->W * 4, h->H, cudaMemcpyHostToDevice, st));
        else CK(cudaMemcpy2DAsync(inout, (size_t)h->W * 4, p, (size_t)h->Wp * 4, (size_t)h->W * 4, h->H, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
    } else if(n == "rawmask") {
        REQUIRE(bytes == npx && !set, "rawmask is read-only, W*H bytes");
        std::vector<uint32_t> bits((size_t)h->H * h->WW);
        d2h(st, bits.data(), h->raw_bits, bits.size() * 4);
        uint8_t* o = (uint8_t*)inout;
        for(int y = 0; y < h->H; ++y) for(int x = 0; x < h->W; ++x) o[(size_t)y * h->W + x] = ((bits[(size_t)y * h->WW + (x >> 5)] >> (x & 31)) & 1u) ? 255 : 0;
    } else if(n == "lastgrad") {
        REQUIRE(bytes == npx * h->MC && !set, "lastgrad is read-only, W*H*C bytes");
        std::vector<uint8_t> g(h->plane * h->grad_bytes());
        d2h(st, g.data(), h->grad[h->frame & 1u], g.size());
        uint8_t* o = (uint8_t*)inout;
        for(int y = 0; y < h->H; ++y) for(int x = 0; x < h->W; ++x) for(int c = 0; c < h->MC; ++c)
            o[((size_t)y * h->W + x) * h->MC + c] = g[((size_t)y * h->Wp + x) * h->grad_bytes() + c];
    } else if(n == "scalars") {
        REQUIRE(bytes == 2 * sizeof(double), "size mismatch for state buffer scalars");
        PbasCtl c0{};
        d2h(st, &c0, h->ctl, sizeof(c0));
        double* d = (double*)inout;
        if(set) {
            const uint32_t f = (uint32_t)d[0];
            if((f ^ h->frame) & 1u) { std::swap(h->grad[0], h->grad[1]); std::swap(h->col[0], h->col[1]); std::swap(h->intents[0], h->intents[1]); } // "latest" planes stay latest
            h->frame = f; c0.former = (float)d[1]; h2d(st, h->ctl, &c0, sizeof(c0));
        }
        else { d[0] = (double)h->frame; d[1] = (double)c0.former; }
    } else REQUIRE(false, "unknown state buffer: " + n);
    LVB_CATCH
}

int lvb_pbas_set_collect_stats(lvb_pbas_handle h, int enabled) {
    LVB_TRY
    REQUIRE(h != nullptr, "null handle");
    h->collect_stats = enabled != 0;
    LVB_CATCH
}
int lvb_pbas_get_stats(lvb_pbas_handle h, uint64_t out[5]) {
    LVB_TRY
    REQUIRE(h != nullptr && h->initialized && out, "algo must be initialized first");
    CK(cudaSetDevice(h->device));
    unsigned long long s[3];
    CK(cudaMemcpyAsync(s, h->d_stats, sizeof(s), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    out[0] = (uint64_t)h->W * h->H * h->stat_frames; out[1] = s[0]; out[2] = s[1]; out[3] = s[2]; out[4] = h->stat_frames;
    LVB_CATCH
}
int lvb_pbas_set_profile(lvb_pbas_handle h, int enabled) {
    LVB_TRY
    REQUIRE(h != nullptr, "null handle");
    h->profile = enabled != 0;
    LVB_CATCH
}
int lvb_pbas_get_profile(lvb_pbas_handle h, double* ms_total, uint64_t* launches) {
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
void* lvb_pbas_stream(lvb_pbas_handle h) { return h ? (void*)h->stream : nullptr; }

} // extern "C"
