// ORACLE — TEST INFRASTRUCTURE ONLY (see lvo_common.hpp header).
// CPU restatement of BackgroundSubtractorPBAS_1ch / _3ch (reference video/src/BackgroundSubtractorPBAS.cpp,
// video/include/litiv/video/BackgroundSubtractorPBAS.hpp; compile-time switches as shipped: SELF_DIFFUSION 1, R2_ACCELERATION 0,
// ADVANCED_MORPH_OPS 0, SC_THRS_VALIDATION 0). Parity pinned: MODE_REFERENCE equals the reference's own BackgroundSubtractorPBAS.cpp
// (oracle/_ref) bit for bit: masks, colour / gradient models, R(x), T(x), mean-min-distance maps (tests/test_ref_pin_cpu.py); the OpenCV
// calls of the gradient image (GaussianBlur 3x3, Scharr, convertScaleAbs, addWeighted) are restated in integers and pinned against
// cv2 4.13 by tests/test_pbas_oracle_cpu.py.
//
// MODE_REFERENCE  the reference's raster loop, glibc rand() clone, float accumulation of the frame's gradient distances in raster order.
// MODE_SNAPSHOT   the deterministic parallel semantics of the GPU: pixels scan frame-start samples; own-slot writes apply at once;
//                 the neighbour ("self-diffusion") writes are queued and applied after the pixel pass; one Philox block per pixel
//                 (sites as in the ViBe / LOBSTER oracles); the 3-channel frame sum of gradient distances is accumulated in 2^-16
//                 fixed point (order-independent) instead of a raster-order float sum.
#pragma once
#include "lvo_common.hpp"
#include "lvo_subsense.hpp"

namespace lvo {

/// cv::borderInterpolate(p, len, BORDER_REFLECT_101)
inline int reflect101(int p, int len) {
    if(len == 1) return 0;
    while(p < 0 || p >= len) { if(p < 0) p = -p; if(p >= len) p = 2 * len - 2 - p; }
    return p;
}

/// PBAS.cpp:80-89 / :125-134: GaussianBlur(3x3, sigma 0 -> [1 2 1]/4, OpenCV's fixed-point path for 8-bit images: round half up),
/// Scharr dx / dy in 16S, convertScaleAbs (|v| saturated to 255), addWeighted(0.5, 0.5) (float, cvRound = half to even). All
/// BORDER_DEFAULT (reflect 101), each stage on the previous stage's full image.
inline void pbas_gradient_image(const uchar* img, int W, int H, int C, uchar* grad) {
    std::vector<uchar> bl((size_t)W * H * C);
    for(int y = 0; y < H; ++y) for(int x = 0; x < W; ++x) for(int c = 0; c < C; ++c) {
        static const int w3[3] = {1, 2, 1};
        int s = 0;
        for(int dy = -1; dy <= 1; ++dy) for(int dx = -1; dx <= 1; ++dx)
            s += w3[dy + 1] * w3[dx + 1] * img[((size_t)reflect101(y + dy, H) * W + reflect101(x + dx, W)) * C + c];
        bl[((size_t)y * W + x) * C + c] = (uchar)((s + 8) >> 4);
    }
    auto B = [&](int y, int x, int c) { return (int)bl[((size_t)reflect101(y, H) * W + reflect101(x, W)) * C + c]; };
    for(int y = 0; y < H; ++y) for(int x = 0; x < W; ++x) for(int c = 0; c < C; ++c) {
        const int gx = 3 * (B(y - 1, x + 1, c) - B(y - 1, x - 1, c)) + 10 * (B(y, x + 1, c) - B(y, x - 1, c)) + 3 * (B(y + 1, x + 1, c) - B(y + 1, x - 1, c));
        const int gy = 3 * (B(y + 1, x - 1, c) - B(y - 1, x - 1, c)) + 10 * (B(y + 1, x, c) - B(y - 1, x, c)) + 3 * (B(y + 1, x + 1, c) - B(y - 1, x + 1, c));
        const int t = std::min(std::abs(gx), 255) + std::min(std::abs(gy), 255);
        grad[((size_t)y * W + x) * C + c] = (uchar)((t >> 1) + ((t & 1) & ((t >> 1) & 1)));
    }
}

struct PBAS {
    // PBAS.hpp:47-54 defaults: R0 = 30, T0 = 16, N = 35, #min = 2
    int color_dist_threshold = 30, n_samples = 35, n_required = 2;
    float default_update_rate = 16.0f;
    int model_channels = 3;
    Mode mode = MODE_REFERENCE;
    uint64_t seed = 0;
    GlibcRand grand;
    int W = 0, H = 0;
    size_t npx = 0, frame_idx = 0;
    bool initialized = false;
    float former_mean_grad_dist = 20.0f; // PBAS.cpp:29
    std::vector<uchar> bg_color, bg_grad; // [N][H][W][C]
    std::vector<float> R, T, meanmin;     // m_oDistThresholdFrame, m_oUpdateRateFrame, m_oMeanMinDistFrame
    std::vector<uchar> raw_mask, last_grad;
    Stats stats;

    uchar* bgc(int s) { return bg_color.data() + (size_t)s * npx * model_channels; }
    uchar* bgg(int s) { return bg_grad.data() + (size_t)s * npx * model_channels; }

    std::vector<uchar> to_model_channels(const uchar* img, int c_in) const {
        if(c_in != 1 && c_in != 3) throw std::runtime_error("input image type must be 8UC1 or 8UC3");
        if(model_channels == 1 && c_in != 1) throw std::runtime_error("input image type must be 8UC1"); // PBAS.cpp:64, :116
        std::vector<uchar> out(npx * model_channels);
        if(c_in == model_channels) std::memcpy(out.data(), img, out.size());
        else for(size_t p = 0; p < npx; ++p) out[p * 3] = out[p * 3 + 1] = out[p * 3 + 2] = img[p];
        return out;
    }

    /// PBAS.cpp:60-110 (1ch) / :284-326 (3ch)
    void initialize(const uchar* img, int w, int h, int c_in) {
        if(!img || w <= 0 || h <= 0) throw std::runtime_error("provided image for initialization must be non-empty and continuous");
        if(n_samples <= 0 || n_required > n_samples) throw std::runtime_error("algo cannot require more sample matches than sample count in model");
        if(!(default_update_rate > 0 && default_update_rate <= 255)) throw std::runtime_error("default update rate must be in ]0,255]"); // PBAS.cpp:32
        W = w; H = h; npx = (size_t)w * h;
        const int C = model_channels;
        const std::vector<uchar> im = to_model_channels(img, c_in);
        std::vector<uchar> grad(npx * C);
        pbas_gradient_image(im.data(), W, H, C, grad.data());
        R.assign(npx, 1.0f); T.assign(npx, default_update_rate); meanmin.assign(npx, 0.0f);
        raw_mask.assign(npx, 0); last_grad = grad;
        bg_color.assign((size_t)n_samples * npx * C, 0); bg_grad.assign((size_t)n_samples * npx * C, 0);
        for(int s = 0; s < n_samples; ++s)
            for(size_t p = 0; p < npx; ++p) {
                const int rnd = mode == MODE_REFERENCE ? grand.next() : philox_draw(seed, 0, (uint32_t)p, (uint32_t)s, DOM_REFRESH);
                int sx, sy;
                sample_pos_7x7(rnd, sx, sy, (int)(p % W), (int)(p / W), 0, W, H);
                const size_t q = (size_t)sy * W + sx;
                for(int c = 0; c < C; ++c) { bgc(s)[p * C + c] = im[q * C + c]; bgg(s)[p * C + c] = grad[q * C + c]; }
            }
        former_mean_grad_dist = 20.0f;
        frame_idx = 0;
        stats = Stats();
        initialized = true;
    }

    /// lv::L2dist<3,uchar> (utils/math.hpp:391-397): squares accumulated in uint16 (wraps mod 65536), float sqrt
    static float l2dist3(const uchar* a, const uchar* b) {
        uint16_t acc = 0;
        for(int c = 0; c < 3; ++c) { const int d = (int)a[c] - (int)b[c]; acc = (uint16_t)(acc + (uint16_t)(d * d)); }
        return (float)std::sqrt((float)acc);
    }

    /// PBAS.cpp:112-271 (1ch) / :328-496 (3ch)
    void apply(const uchar* img, int c_in, uchar* fgmask, double lr_override) {
        if(!initialized) throw std::runtime_error("algo must be initialized first");
        const int C = model_channels;
        const std::vector<uchar> im = to_model_channels(img, c_in);
        std::vector<uchar> grad(npx * C);
        pbas_gradient_image(im.data(), W, H, C, grad.data());
        last_grad = grad;
        const size_t N = (size_t)n_samples, REQ = (size_t)n_required;
        ++frame_idx; // Philox counter only
        const uint32_t fr = (uint32_t)frame_idx;
        std::fill(raw_mask.begin(), raw_mask.end(), 0);
        size_t tot_grad_int = 0, bad = 1;       // 1ch: nFrameTotGradDist (size_t), nFrameTotBadSamplesCount = 1 (:139-140)
        float tot_grad_flt = 0; uint64_t tot_grad_fix = 0; // 3ch: fFrameTotGradDist (float, :354) / snapshot mode: 2^-16 fixed point
        const float grad_w = 10.0f / former_mean_grad_dist; // BGSPBAS_GRAD_WEIGHT_ALPHA / m_fFormerMeanGradDist
        std::vector<std::pair<size_t, int>> queue; // (target pixel, slot)
        for(size_t p = 0; p < npx; ++p) {
            const int x = (int)(p % W), y = (int)(p / W);
            const uchar* cur = im.data() + p * C;
            const uchar* cg = grad.data() + p * C;
            auto draw = [&](uint32_t site) -> size_t {
                return (size_t)(mode == MODE_REFERENCE ? grand.next() : philox_draw(seed, fr, (uint32_t)p, site, DOM_APPLY));
            };
            float min_dist = 255.0f;
            const float thr = R[p] * (float)(size_t)color_dist_threshold;
            size_t good = 0, s = 0;
            while(good < REQ && s < N) {
                float sum, gd_f = 0; size_t gd_i = 0;
                if(C == 1) {
                    const size_t cd = L1dist_u8(cur[0], bgc((int)s)[p]);
                    gd_i = L1dist_u8(cg[0], bgg((int)s)[p]);
                    sum = std::min((grad_w * (float)gd_i) + (float)cd, 255.0f);
                } else {
                    const float cd = l2dist3(cur, bgc((int)s) + p * 3);
                    gd_f = l2dist3(cg, bgg((int)s) + p * 3);
                    sum = std::min((grad_w * gd_f) + cd, 255.0f);
                }
                if(sum <= thr) { if(min_dist > sum) min_dist = sum; ++good; }
                else {
                    if(C == 1) tot_grad_int += gd_i;
                    else if(mode == MODE_REFERENCE) tot_grad_flt += gd_f;
                    else tot_grad_fix += (uint64_t)std::llrintf(gd_f * 65536.0f);
                    ++bad;
                }
                ++s;
            }
            stats.samples_scanned += s;
            meanmin[p] = (meanmin[p] * (float)(N - 1) + (min_dist / 255.0f)) / (float)N;
            if(good < REQ) {
                raw_mask[p] = 255; ++stats.fg_px;
                T[p] += 1.0f / (meanmin[p] * 255.0f + 1.0f);
                if(T[p] > 200.0f) T[p] = 200.0f;
            } else {
                const double ce = lr_override > 0 ? std::ceil(lr_override) : std::ceil((double)T[p]);
                const size_t LR = std::isinf(ce) ? SIZE_MAX : (size_t)ce;
                if((draw(0) % LR) == 0) {
                    const size_t slot = draw(1) % N;
                    for(int c = 0; c < C; ++c) { bgc((int)slot)[p * C + c] = cur[c]; bgg((int)slot)[p * C + c] = cg[c]; }
                    ++stats.sample_writes;
                }
                if((draw(2) % LR) == 0) {
                    int nx, ny;
                    neighbor_pos_3x3((int)draw(3), nx, ny, x, y, 0, W, H);
                    const size_t slot = mode == MODE_REFERENCE ? draw(4) % N : (draw(1) / N) % N;
                    const size_t q = (size_t)ny * W + nx;
                    // BGSPBAS_USE_SELF_DIFFUSION: the neighbour receives ITS OWN current colour / gradient (:190-191, :424-425)
                    if(mode == MODE_REFERENCE) {
                        for(int c = 0; c < C; ++c) { bgc((int)slot)[q * C + c] = im[q * C + c]; bgg((int)slot)[q * C + c] = grad[q * C + c]; }
                        ++stats.sample_writes;
                    } else queue.emplace_back(q, (int)slot);
                }
                T[p] -= 0.05f / (meanmin[p] * 255.0f + 1.0f);
                if(T[p] < 2.0f) T[p] = 2.0f;
            }
            if(R[p] < 0.6f + meanmin[p] * 5.0f + 0.0f) { if(R[p] < 99.0f) R[p] *= 1.05f; }
            else if(R[p] > 0.6f) R[p] *= 0.95f;
        }
        for(const auto& w : queue) {
            for(int c = 0; c < C; ++c) { bgc(w.second)[w.first * C + c] = im[w.first * C + c]; bgg(w.second)[w.first * C + c] = grad[w.first * C + c]; }
            ++stats.sample_writes;
        }
        float tot;
        if(C == 1) tot = (float)tot_grad_int;
        else tot = mode == MODE_REFERENCE ? tot_grad_flt : (float)((double)tot_grad_fix * (1.0 / 65536.0));
        former_mean_grad_dist = std::max(tot / (float)bad, 20.0f); // :224 / :456
        stats.roi_px += npx; ++stats.frames;
        median_binary(raw_mask.data(), fgmask, W, H, 9);          // :269 / :494 (ADVANCED_MORPH_OPS 0)
    }

    /// PBAS.cpp:37-54
    void get_background_image(uchar* out) const {
        if(!initialized) throw std::runtime_error("algo must be initialized first");
        const size_t n = npx * model_channels;
        std::vector<float> acc(n, 0.f);
        for(int s = 0; s < n_samples; ++s) {
            const uchar* b = bg_color.data() + (size_t)s * n;
            for(size_t i = 0; i < n; ++i) acc[i] += ((float)b[i]) / n_samples;
        }
        for(size_t i = 0; i < n; ++i) out[i] = sat_u8(acc[i]);
    }
};

} // namespace lvo
