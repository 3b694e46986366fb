// litiv_b200 — host side of the LBSP edge detector entry points (include/litiv_b200.h: lvb_edge_*). Included at the end of
// litiv_b200.cu (same translation unit: CK / REQUIRE / LVB_TRY, dalloc, make_image_tmap come from there).
// Mirrors EdgeDetectorLBSP (imgproc/include/litiv/imgproc/EdgeDetectorLBSP.hpp:33-83): a detector object with a level count and a
// hysteresis factor, apply_threshold(img, mask, thr) and apply(img, confidence); like the reference object it keeps its gradient map
// and edge mask between calls (the two mask rows the suppression never writes are observable state, see edge_px.cuh).
// Deviation: when the image size changes between calls the persisted rows restart from zero (the reference's std::vector::resize keeps
// the old bytes at their linear offsets, which then land on unrelated pixels).
#pragma once
#include "edge.cuh"

struct lvb_edge_context {
    int device = 0, levels = 3; double hyst = 0.5; bool normalize = false;   // m_bNormalizeOutput
    int W = 0, H = 0, C = 0;
    cudaStream_t stream = nullptr;
    std::vector<int> Wl, Hl; std::vector<size_t> pitch;
    std::vector<uint8_t*> img; std::vector<uchar4*> V; std::vector<CUtensorMap> tmap; std::vector<int> use_tma;
    uint8_t *mask = nullptr, *out = nullptr; int* flag = nullptr; unsigned* minmax = nullptr;
    // hysteresis by union-find on row runs (edge.cuh): bit planes, parents, per-word run ranks
    uint32_t *uf_E = nullptr, *uf_S = nullptr, *uf_parent = nullptr; uint16_t* uf_rank = nullptr; int WW = 0, RS = 0;
    bool use_sweeps = getenv("LVB_EDGE_SWEEPS") != nullptr;   // debugging aid: the relaxation sweeps of round 1 instead
    uint64_t flood_sweeps = 0;   // relaxation sweeps of the latest call (diagnostic)

    void free_all() {
        for(uint8_t* p : img) if(p) cudaFree(p);
        for(uchar4* p : V) if(p) cudaFree(p);
        if(mask) cudaFree(mask);
        if(out) cudaFree(out);
        if(flag) cudaFree(flag);
        if(minmax) cudaFree(minmax);
        for(void* p : {(void*)uf_E, (void*)uf_S, (void*)uf_parent, (void*)uf_rank}) if(p) cudaFree(p);
        uf_E = uf_S = uf_parent = nullptr; uf_rank = nullptr;
        img.clear(); V.clear(); tmap.clear(); use_tma.clear(); Wl.clear(); Hl.clear(); pitch.clear();
        mask = out = nullptr; flag = nullptr; minmax = nullptr; W = H = C = 0;
    }
};

namespace {

void edge_prepare(lvb_edge_context* c, const uint8_t* src, int W, int H, int C, size_t src_step = 0, cudaMemcpyKind kind = cudaMemcpyHostToDevice) {
    REQUIRE(src && C >= 1 && C <= 4, "input image must be non-empty and continuous, 8UC1 .. 8UC4");   // EdgeDetectorLBSP.cpp:144-160, 392-393
    std::vector<int> Wl(1, W), Hl(1, H);
    for(int l = 1; l < c->levels; ++l) { Wl.push_back((Wl.back() + 1) / 2); Hl.push_back((Hl.back() + 1) / 2); }
    REQUIRE(Wl.back() >= 5 && Hl.back() >= 5, "image too small for the number of pyramid levels");
    CK(cudaSetDevice(c->device));
    if(W != c->W || H != c->H || C != c->C) {
        CK(cudaStreamSynchronize(c->stream));
        c->free_all();
        c->Wl = Wl; c->Hl = Hl;
        c->tmap.resize(c->levels); c->use_tma.assign(c->levels, 0);
        for(int l = 0; l < c->levels; ++l) {
            c->pitch.push_back(((size_t)Wl[l] * C + 127) / 128 * 128);
            c->img.push_back(dalloc<uint8_t>(c->stream, c->pitch[l] * Hl[l]));
            c->V.push_back(dalloc<uchar4>(c->stream, (size_t)Wl[l] * Hl[l]));
            c->use_tma[l] = make_image_tmap(&c->tmap[l], c->img[l], Wl[l], Hl[l], C, c->pitch[l]) ? 1 : 0;
        }
        c->mask = dalloc<uint8_t>(c->stream, (size_t)W * H);   // zero = "may belong to an edge", the fresh vector of the reference
        c->out = dalloc<uint8_t>(c->stream, (size_t)W * H);
        c->flag = dalloc<int>(c->stream, 1);
        c->minmax = dalloc<unsigned>(c->stream, 2);
        c->WW = (W + 31) / 32; c->RS = (W + 1) / 2 + 1;   // at most ceil(W/2) runs per row
        c->uf_E = dalloc<uint32_t>(c->stream, (size_t)H * c->WW); c->uf_S = dalloc<uint32_t>(c->stream, (size_t)H * c->WW);
        c->uf_parent = dalloc<uint32_t>(c->stream, 1 + (size_t)H * c->RS); c->uf_rank = dalloc<uint16_t>(c->stream, (size_t)H * c->WW);
        c->W = W; c->H = H; c->C = C;   // last: a failed allocation leaves W == 0, so the next call starts over instead of using half a set of maps
    }
    cudaStream_t st = c->stream;
    CK(cudaMemcpy2DAsync(c->img[0], c->pitch[0], src, src_step ? src_step : (size_t)W * C, (size_t)W * C, H, kind, st));
    const dim3 b(32, 8);
    for(int l = 0; l + 1 < c->levels; ++l) {   // apply_internal_lookup: the pyramid
        const dim3 g((Wl[l + 1] * C + 31) / 32, (Hl[l + 1] + 7) / 8);
        lvb_edge::edge_pyr_down_kernel<<<g, b, 0, st>>>(c->img[l], c->pitch[l], Wl[l], Hl[l], C, c->img[l + 1], c->pitch[l + 1], Wl[l + 1], Hl[l + 1]);
        LAUNCHED();
    }
    for(int l = c->levels - 1; l >= 0; --l) {  // per-level LBSP gradient with the min-|.| combination with the coarser level in its epilogue
        LbspGradArgs A{};
        A.W = Wl[l]; A.H = Hl[l]; A.img = c->img[l]; A.ipitch = c->pitch[l]; A.out = c->V[l]; A.use_tma = c->use_tma[l];
        A.combine = 1; A.coarse = l + 1 < c->levels ? c->V[l + 1] : nullptr; A.Wc = l + 1 < c->levels ? Wl[l + 1] : 0;
        const dim3 gg((Wl[l] + TILE_W - 1) / TILE_W, (Hl[l] + TILE_H - 1) / TILE_H), gb(TILE_W, TILE_H);
        if(C == 1) lbsp_gradient_kernel<1><<<gg, gb, 0, st>>>(A, c->tmap[l]); else if(C == 2) lbsp_gradient_kernel<2><<<gg, gb, 0, st>>>(A, c->tmap[l]);
        else if(C == 3) lbsp_gradient_kernel<3><<<gg, gb, 0, st>>>(A, c->tmap[l]); else lbsp_gradient_kernel<4><<<gg, gb, 0, st>>>(A, c->tmap[l]);
        LAUNCHED();
    }
}

/// apply_internal_threshold (:166-375) for one detection threshold, on the maps edge_prepare left on the device
void edge_pass(lvb_edge_context* c, unsigned hi, int accumulate) {
    const unsigned lo = (unsigned)(uint8_t)(hi * c->hyst);
    cudaStream_t st = c->stream;
    const int W = c->W, H = c->H;
    lvb_edge::EdgeMaps m{};
    m.V0 = c->V[0]; m.W = W; m.H = H;
    if(c->levels > 1) { m.V1 = c->V[1]; m.W1 = c->Wl[1]; m.H1 = c->Hl[1]; }
    const dim3 b(32, 8), g((W + 31) / 32, (H + 7) / 8);
    lvb_edge::edge_nms_kernel<<<g, b, 0, st>>>(m, lo, hi, c->mask, c->use_sweeps ? nullptr : c->uf_E, c->use_sweeps ? nullptr : c->uf_S, c->WW); LAUNCHED();
    if(!c->use_sweeps) {   // hysteresis (:353-372) as connected components of {maybe, edge} that hold an edge: no host round trip
        lvb_edge::EdgeUF U{};
        U.W = W; U.H = H; U.WW = c->WW; U.RS = c->RS; U.E = c->uf_E; U.S = c->uf_S; U.parent = c->uf_parent; U.rankbase = c->uf_rank; U.mask = c->mask;
        const int rb = (H + 7) / 8;
        lvb_edge::edge_uf_init<<<rb, 256, 0, st>>>(U); LAUNCHED();
        lvb_edge::edge_uf_union<<<rb, 256, 0, st>>>(U); LAUNCHED();
        lvb_edge::edge_uf_apply<<<rb, 256, 0, st>>>(U); LAUNCHED();
    } else {
        const dim3 fb(lvb_edge::FL_W, lvb_edge::FL_H), fg((W + lvb_edge::FL_W - 1) / lvb_edge::FL_W, (H + lvb_edge::FL_H - 1) / lvb_edge::FL_H);
        constexpr int SWEEPS_PER_CHECK = 4;
        const uint64_t cap = (uint64_t)W * H + SWEEPS_PER_CHECK;   // every sweep but the last turns at least one pixel
        for(uint64_t done = 0;;) {
            // only the last sweep of a group reports: a group ends the loop when its last sweep found nothing left to turn
            for(int i = 0; i < SWEEPS_PER_CHECK; ++i) {
                if(i == SWEEPS_PER_CHECK - 1) CK(cudaMemsetAsync(c->flag, 0, sizeof(int), st));
                lvb_edge::edge_flood_kernel<<<fg, fb, 0, st>>>(c->mask, W, H, c->flag); LAUNCHED();
            }
            done += SWEEPS_PER_CHECK; c->flood_sweeps += SWEEPS_PER_CHECK;
            int changed = 0;
            CK(cudaMemcpyAsync(&changed, c->flag, sizeof(int), cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            if(!changed) break;
            REQUIRE(done < cap, "edge hysteresis did not converge");
        }
    }
    lvb_edge::edge_output_kernel<<<g, b, 0, st>>>(c->mask, W, H, c->out, accumulate); LAUNCHED();
}

} // namespace

extern "C" {

int lvb_edge_create(int levels, double hyst_low_factor, int device, lvb_edge_handle* out) {
    LVB_TRY
    REQUIRE(out != nullptr, "null output handle");
    REQUIRE(levels >= 1, "number of pyramid levels must be positive");                                                       // EdgeDetectorLBSP.cpp:31
    REQUIRE(hyst_low_factor > 0 && hyst_low_factor < 1, "lower hysteresis threshold factor must be between 0 and 1");       // :32
    const int ndev = lvb_device_count();
    REQUIRE(ndev > 0, "no CUDA device available: litiv_b200 has no CPU fallback");
    REQUIRE(device >= 0 && device < ndev, "invalid CUDA device id");
    CK(cudaSetDevice(device));
    lvb_edge_context* c = new lvb_edge_context();
    c->device = device; c->levels = levels; c->hyst = hyst_low_factor;
    try { CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)); } catch(...) { delete c; throw; }
    *out = c;
    LVB_CATCH
}

int lvb_edge_destroy(lvb_edge_handle h) {
    if(!h) return 0;
    cudaSetDevice(h->device);
    if(h->stream) cudaStreamSynchronize(h->stream);
    h->free_all();
    if(h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return 0;
}

int lvb_edge_set_normalize(lvb_edge_handle h, int normalize_output) {
    LVB_TRY
    REQUIRE(h != nullptr, "null argument");
    h->normalize = normalize_output != 0;
    LVB_CATCH
}

double lvb_edge_default_threshold(void) { return 8.0 / 16.0; }   // EDGLBSP_DEFAULT_DET_THRESHOLD (EdgeDetectorLBSP.hpp:30)

int lvb_edge_apply_threshold(lvb_edge_handle h, const uint8_t* img, int W, int H, int C, uint8_t* edges, double threshold) {
    LVB_TRY
    REQUIRE(h != nullptr && edges != nullptr, "null argument");
    if(threshold < 0 || threshold > 1) threshold = lvb_edge_default_threshold();   // :394-395
    h->flood_sweeps = 0;
    edge_prepare(h, img, W, H, C);
    edge_pass(h, (unsigned)(uint8_t)(threshold * 16), 0);
    CK(cudaMemcpyAsync(edges, h->out, (size_t)W * H, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    LVB_CATCH
}

int lvb_edge_apply(lvb_edge_handle h, const uint8_t* img, int W, int H, int C, uint8_t* confidence) {
    LVB_TRY
    REQUIRE(h != nullptr && confidence != nullptr, "null argument");
    h->flood_sweeps = 0;
    edge_prepare(h, img, W, H, C);
    CK(cudaMemsetAsync(h->out, 0, (size_t)W * H, h->stream));
    for(unsigned t = 0; t < 16; ++t) edge_pass(h, t, 1);   // :418-430: the gradient map does not depend on the threshold
    if(h->normalize) {                                     // :431-432
        const unsigned init[2] = {255u, 0u};
        CK(cudaMemcpyAsync(h->minmax, init, sizeof(init), cudaMemcpyHostToDevice, h->stream));
        const size_t n = (size_t)W * H;
        const int grid = (int)std::min<size_t>((n + 255) / 256, 148 * 8);
        lvb_edge::edge_minmax_kernel<<<grid, 256, 0, h->stream>>>(h->out, n, h->minmax); LAUNCHED();
        lvb_edge::edge_normalize_kernel<<<grid, 256, 0, h->stream>>>(h->out, n, h->minmax); LAUNCHED();
    }
    CK(cudaMemcpyAsync(confidence, h->out, (size_t)W * H, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    LVB_CATCH
}

int lvb_edge_get_gradient_map(lvb_edge_handle h, uint8_t* out) {
    LVB_TRY
    REQUIRE(h != nullptr && out != nullptr, "null argument");
    REQUIRE(h->W > 0, "no pass has run yet");
    CK(cudaSetDevice(h->device));
    CK(cudaMemcpyAsync(out, h->V[0], (size_t)h->W * h->H * 4, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    LVB_CATCH
}

uint64_t lvb_edge_flood_sweeps(lvb_edge_handle h) { return h ? h->flood_sweeps : 0; }

/* the same pass on a frame that already lives in device memory (row pitch d_step bytes); d_edges_or_null: W*H device bytes. The call
 * returns when the result is complete (the hysteresis loop reads a convergence flag back). Written at the end of round 1 and NOT yet
 * run on a device: tools/bench_edge.py is its first user. */
int lvb_edge_apply_threshold_device(lvb_edge_handle h, const uint8_t* d_img, int W, int H, int C, size_t d_step, uint8_t* d_edges_or_null, double threshold) {
    LVB_TRY
    REQUIRE(h != nullptr, "null argument");
    REQUIRE(d_step >= (size_t)W * C, "row pitch smaller than a row");
    if(threshold < 0 || threshold > 1) threshold = lvb_edge_default_threshold();
    h->flood_sweeps = 0;
    edge_prepare(h, d_img, W, H, C, d_step, cudaMemcpyDeviceToDevice);
    edge_pass(h, (unsigned)(uint8_t)(threshold * 16), 0);
    if(d_edges_or_null) CK(cudaMemcpyAsync(d_edges_or_null, h->out, (size_t)W * H, cudaMemcpyDeviceToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    LVB_CATCH
}
void* lvb_edge_stream(lvb_edge_handle h) { return h ? (void*)h->stream : nullptr; }


} // extern "C"
