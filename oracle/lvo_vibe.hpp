// ORACLE — TEST INFRASTRUCTURE ONLY (see lvo_common.hpp header).
// CPU restatement of BackgroundSubtractorViBe_1ch / _3ch (reference video/src/BackgroundSubtractorViBe.cpp,
// video/include/litiv/video/BackgroundSubtractorViBe.hpp), the colour-only ancestor of LOBSTER's sample-consensus scan.
// Parity pinned: MODE_REFERENCE equals the reference's own BackgroundSubtractorViBe.cpp (oracle/_ref) bit for bit, masks and models
// (tests/test_ref_pin_cpu.py); the reference holds no test, golden vector or fixture for ViBe.
//
// Two modes, like the other oracles:
//   MODE_REFERENCE  the reference's raster loop with the glibc rand() clone (same draw order as the source);
//   MODE_SNAPSHOT   the deterministic parallel semantics the GPU implements: every pixel scans frame-start samples, own-slot
//                   writes apply at once, neighbour writes are queued and applied after the pixel pass in raster order of
//                   their SOURCE pixel (last writer wins), draws are one Philox block per pixel (sites 0..3: own decision,
//                   own slot | neighbour slot = (draw1 / N) % N, neighbour decision, neighbour position).
#pragma once
#include "lvo_common.hpp"
#include "lvo_subsense.hpp"

namespace lvo {

struct ViBe {
    // ViBe.hpp:41-47 defaults: colour distance threshold 20, N = 20, #min = 2, learning rate ("subsampling factor") 16
    int color_dist_threshold = 20, n_samples = 20, n_required = 2;
    int model_channels = 3; // BackgroundSubtractorViBe_1ch (8UC1 only) or _3ch (8UC3, or 8UC1 expanded by cvtColor GRAY2BGR)
    Mode mode = MODE_REFERENCE;
    uint64_t seed = 0;
    GlibcRand grand;
    int W = 0, H = 0;
    size_t npx = 0, frame_idx = 0;
    bool initialized = false;
    std::vector<uchar> bg; // [N][H][W][C]
    Stats stats;

    uchar* sample(int s) { return bg.data() + (size_t)s * npx * model_channels; }

    /// cvtColor(GRAY2BGR) of ViBe.cpp:121-124 / :147-150 when a one-channel image is given to the three-channel class
    std::vector<uchar> to_model_channels(const uchar* img, int c_in) const {
        if(c_in != 1 && c_in != 3) throw std::runtime_error("input image type must be 8UC1 or 8UC3");
        if(model_channels == 1 && c_in != 1) throw std::runtime_error("input image type must be 8UC1"); // ViBe.cpp:61, :83
        std::vector<uchar> out(npx * model_channels);
        if(c_in == model_channels) std::memcpy(out.data(), img, out.size());
        else for(size_t p = 0; p < npx; ++p) out[p * 3] = out[p * 3 + 1] = out[p * 3 + 2] = img[p];
        return out;
    }

    /// ViBe.cpp:58-76 (1ch) / :115-138 (3ch): every sample of every pixel drawn from the 7x7 neighbourhood (border 0)
    void initialize(const uchar* img, int w, int h, int c_in) {
        if(!img || w <= 0 || h <= 0) throw std::runtime_error("provided image for initialization must be non-empty and continuous");
        if(n_samples <= 0 || n_required > n_samples) throw std::runtime_error("algo cannot require more sample matches than sample count in model");
        W = w; H = h; npx = (size_t)w * h;
        const int C = model_channels;
        const std::vector<uchar> im = to_model_channels(img, c_in);
        bg.assign((size_t)n_samples * npx * C, 0);
        for(int s = 0; s < n_samples; ++s)
            for(size_t p = 0; p < npx; ++p) {
                const int ox = (int)(p % W), oy = (int)(p / W);
                const int rnd = mode == MODE_REFERENCE ? grand.next() : philox_draw(seed, 0, (uint32_t)p, (uint32_t)s, DOM_REFRESH);
                int sx, sy;
                sample_pos_7x7(rnd, sx, sy, ox, oy, 0, W, H);
                for(int c = 0; c < C; ++c) sample(s)[p * C + c] = im[((size_t)sy * W + sx) * C + c];
            }
        frame_idx = 0;
        stats = Stats();
        initialized = true;
    }

    /// the match test of one sample: ViBe.cpp:93 (1ch: L1 < thr) / :167-171 (3ch: L2dist < thr*3, BGSVIBE_USE_L1_DISTANCE_CHECK 0,
    /// BGSVIBE_USE_SC_THRS_VALIDATION 0). lv::L2dist<3,uchar> (utils/math.hpp:391-397) accumulates the squared differences in
    /// decltype(L2sqrdist(uchar,uchar)) = uint16 (math.hpp:301-306), so the sum wraps mod 65536 before the float sqrt.
    bool matches(const uchar* cur, const uchar* b) const {
        if(model_channels == 1) return (size_t)L1dist_u8(cur[0], b[0]) < (size_t)color_dist_threshold;
        uint16_t acc = 0;
        for(int c = 0; c < 3; ++c) { const int d = (int)cur[c] - (int)b[c]; acc = (uint16_t)(acc + (uint16_t)(d * d)); }
        return (float)std::sqrt((float)acc) < (float)((size_t)color_dist_threshold * 3);
    }

    struct NbWrite { size_t target; int slot; uchar col[3]; };

    /// ViBe.cpp:78-110 (1ch) / :140-191 (3ch)
    void apply(const uchar* img, int c_in, uchar* fgmask, double lr) {
        if(!initialized) throw std::runtime_error("algo must be initialized first");
        if(!(lr > 0)) throw std::runtime_error("learning rate must be a positive value");
        const int C = model_channels;
        const std::vector<uchar> im = to_model_channels(img, c_in);
        const size_t N = (size_t)n_samples, REQ = (size_t)n_required;
        const size_t LR = std::isinf(lr) ? SIZE_MAX : (size_t)std::ceil(lr);
        ++frame_idx; // Philox counter only
        const uint32_t fr = (uint32_t)frame_idx;
        std::memset(fgmask, 0, npx);
        std::vector<NbWrite> queue;
        for(size_t p = 0; p < npx; ++p) {
            const int x = (int)(p % W), y = (int)(p / W);
            const uchar* cur = im.data() + p * C;
            auto draw = [&](uint32_t site) -> size_t {
                return (size_t)(mode == MODE_REFERENCE ? grand.next() : philox_draw(seed, fr, (uint32_t)p, site, DOM_APPLY));
            };
            size_t good = 0, s = 0;
            while(good < REQ && s < N) {
                if(matches(cur, sample((int)s) + p * C)) ++good;
                ++s;
            }
            stats.samples_scanned += s;
            if(good < REQ) { fgmask[p] = 255; ++stats.fg_px; continue; }
            if((draw(0) % LR) == 0) {
                const size_t slot = draw(1) % N;
                for(int c = 0; c < C; ++c) sample((int)slot)[p * C + c] = cur[c];
                ++stats.sample_writes;
            }
            if((draw(2) % LR) == 0) {
                int nx, ny;
                neighbor_pos_3x3((int)draw(3), nx, ny, x, y, 0, W, H);
                const size_t slot = mode == MODE_REFERENCE ? draw(4) % N : (draw(1) / N) % N;
                const size_t q = (size_t)ny * W + nx;
                if(mode == MODE_REFERENCE) {
                    for(int c = 0; c < C; ++c) sample((int)slot)[q * C + c] = cur[c];
                    ++stats.sample_writes;
                } else {
                    NbWrite w; w.target = q; w.slot = (int)slot;
                    for(int c = 0; c < C; ++c) w.col[c] = cur[c];
                    queue.push_back(w);
                }
            }
        }
        for(const NbWrite& w : queue) {
            for(int c = 0; c < C; ++c) sample(w.slot)[w.target * C + c] = w.col[c];
            ++stats.sample_writes;
        }
        stats.roi_px += npx; ++stats.frames;
    }

    /// ViBe.cpp:32-49: float mean accumulated sample by sample (x / N each), convertTo(CV_8U) = round-half-even + saturate
    void get_background_image(uchar* out) const {
        if(!initialized) throw std::runtime_error("algo must be initialized first");
        const size_t n = npx * model_channels;
        std::vector<float> acc(n, 0.f);
        for(int s = 0; s < n_samples; ++s) {
            const uchar* b = bg.data() + (size_t)s * n;
            for(size_t i = 0; i < n; ++i) acc[i] += ((float)b[i]) / n_samples;
        }
        for(size_t i = 0; i < n; ++i) out[i] = sat_u8(acc[i]);
    }
};

} // namespace lvo
