// TEST INFRASTRUCTURE. CPU emulation of the edge-detector kernels (litiv_b200/csrc/edge.cuh): the SAME per-pixel bodies
// (litiv_b200/csrc/edge_px.cuh, compiled here by g++ through cuda_runtime.h's host definitions of __host__ __device__) driven by plain
// loops in an arbitrary (here: reversed) pixel order, to check on a machine without a GPU that the order-free restatement the kernels
// use equals the sequential oracle (oracle/lvo_edge_lbsp.hpp). The per-level LBSP gradient comes from the oracle's dense primitive, which
// the GPU parity tests already tie to lbsp_gradient_kernel. Not shipped, not linked into the product.
#include "../litiv_b200/csrc/edge_px.cuh"
#include "../oracle/lvo_common.hpp"
#include <vector>
#include <cstring>

using namespace lvb_edge;

struct Emul {
    int levels = 3; double hyst = 0.5;
    int W = 0, H = 0, C = 0;
    std::vector<int> Wl, Hl;
    std::vector<std::vector<uchar>> img;
    std::vector<std::vector<uchar4>> V;
    std::vector<uchar> mask, out;
    long sweeps = 0;

    void prepare(const uchar* src, int w, int h, int c) {
        if(w != W || h != H || c != C) {
            W = w; H = h; C = c;
            Wl.assign(1, w); Hl.assign(1, h);
            for(int l = 1; l < levels; ++l) { Wl.push_back((Wl.back() + 1) / 2); Hl.push_back((Hl.back() + 1) / 2); }
            img.assign(levels, {}); V.assign(levels, {});
            for(int l = 0; l < levels; ++l) { img[l].assign((size_t)Wl[l] * Hl[l] * C, 0); V[l].assign((size_t)Wl[l] * Hl[l], uchar4{0, 0, 0, 0}); }
            mask.assign((size_t)W * H, 0); out.assign((size_t)W * H, 0);
        }
        std::memcpy(img[0].data(), src, (size_t)W * H * C);
        for(int l = 0; l + 1 < levels; ++l)
            for(int y = Hl[l + 1] - 1; y >= 0; --y) for(int xk = Wl[l + 1] * C - 1; xk >= 0; --xk)
                img[l + 1][(size_t)y * Wl[l + 1] * C + xk] = pyr_down_px(img[l].data(), (size_t)Wl[l] * C, Wl[l], Hl[l], C, 2 * y, 2 * (xk / C), xk % C);
        for(int l = levels - 1; l >= 0; --l) {
            lvo::lbsp_gradient_dense(img[l].data(), Wl[l], Hl[l], C, (uchar*)V[l].data());
            for(int y = 0; y < Hl[l]; ++y) for(int x = 0; x < Wl[l]; ++x) {
                const uchar4 c4 = l + 1 < levels ? V[l + 1][(size_t)(y >> 1) * Wl[l + 1] + (x >> 1)] : edge_init_value();
                V[l][(size_t)y * Wl[l] + x] = edge_combine(V[l][(size_t)y * Wl[l] + x], c4);
            }
        }
    }
    void pass(unsigned hi, int accumulate) {
        const unsigned lo = (unsigned)(uchar)(hi * hyst);
        EdgeMaps m{};
        m.V0 = V[0].data(); m.W = W; m.H = H;
        if(levels > 1) { m.V1 = V[1].data(); m.W1 = Wl[1]; m.H1 = Hl[1]; }
        for(int y = H - 3; y >= 0; --y) for(int x = W - 1; x >= 0; --x) mask[(size_t)y * W + x] = edge_mask_value(m, y, x, lo, hi, mask.data());
        for(bool changed = true; changed;) {   // Jacobi sweeps: the slowest legal schedule of edge_flood_kernel
            changed = false; ++sweeps;
            std::vector<uchar> nxt = mask;
            for(int y = 0; y < H; ++y) for(int x = 0; x < W; ++x) {
                if(mask[(size_t)y * W + x] != EDGE_MAYBE) continue;
                bool set = false;
                for(int dy = -1; dy <= 1 && !set; ++dy) for(int dx = -1; dx <= 1; ++dx) {
                    const int yy = y + dy, xx = x + dx;
                    if((dy || dx) && yy >= 0 && xx >= 0 && yy < H && xx < W && mask[(size_t)yy * W + xx] == EDGE_YES) { set = true; break; }
                }
                if(set) { nxt[(size_t)y * W + x] = EDGE_YES; changed = true; }
            }
            mask.swap(nxt);
        }
        for(int y = 0; y < H; ++y) for(int x = 0; x < W; ++x) {   // edge_output_kernel
            const size_t i = (size_t)y * W + x;
            const uchar v = mask[i];
            const bool on = v >= EDGE_YES;
            if(y >= H - 2 && v == EDGE_YES) mask[i] = EDGE_STALE_YES;
            if(accumulate) { const unsigned a = out[i] + (on ? 16u : 0u); out[i] = (uchar)(a > 255u ? 255u : a); }
            else out[i] = on ? 255 : 0;
        }
    }
};

extern "C" {
void* emul_create(int levels, double hyst) { Emul* e = new Emul(); e->levels = levels; e->hyst = hyst; return e; }
void emul_destroy(void* h) { delete (Emul*)h; }
void emul_apply_threshold(void* h, const uchar* img, int W, int H, int C, uchar* out, double thr) {
    Emul* e = (Emul*)h;
    if(thr < 0 || thr > 1) thr = 0.5;
    e->prepare(img, W, H, C);
    e->pass((unsigned)(uchar)(thr * 16), 0);
    std::memcpy(out, e->out.data(), (size_t)W * H);
}
void emul_apply(void* h, const uchar* img, int W, int H, int C, uchar* out) {
    Emul* e = (Emul*)h;
    e->prepare(img, W, H, C);
    std::fill(e->out.begin(), e->out.end(), 0);
    for(unsigned t = 0; t < 16; ++t) e->pass(t, 1);
    std::memcpy(out, e->out.data(), (size_t)W * H);
}
void emul_gradient_map(void* h, uchar* out) { Emul* e = (Emul*)h; std::memcpy(out, e->V[0].data(), (size_t)e->W * e->H * 4); }
}
