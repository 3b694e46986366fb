// ORACLE — TEST INFRASTRUCTURE ONLY (see lvo_common.hpp header).
// CPU restatement of EdgeDetectorLBSP (reference imgproc/src/EdgeDetectorLBSP.cpp:26-417, imgproc/include/litiv/imgproc/
// EdgeDetectorLBSP.hpp; compile-time switches as shipped: USE_5x5_NON_MAX_SUPP 1, USE_MIN_GRAD_ORIENT 1, USE_3_AXIS_ORIENT 1; Gaussian
// sigma 0, i.e. no pre-blur). SURVEY §8f rank 4: the checker of lvb_edge_*  and of the per-pixel
// primitive (lvb_lbsp_gradient). Parity pinned: equals the reference's own imgproc/src/EdgeDetectorLBSP.cpp (oracle/_ref) bit for bit: edge masks,
// confidence maps and the detector's two persistent buffers after every call (tests/test_ref_pin_cpu.py); the reference holds no test for the detector.
//
// The restatement keeps the reference's OBSERVABLE behaviour, including three things that look unintended:
//   * the non-maximum-suppression loop classifies gradient row r+2 into mask row r (:263, :270-272), so the output is shifted up by two rows;
//   * mask rows H and H+1 of the padded map are never written by that loop: they keep what the object's previous call left there
//     (zero = "may belong to an edge" on a fresh object) and are the source of output rows H-2, H-1. In practice they stay empty: the
//     two mask rows above them hold the classes of gradient rows H-2, H-1, which lie in the LBSP border (magnitude 0, never a maximum),
//     so no seed can reach them (asserted in tests/test_edge_emul_cpu.py);
//   * the gradient map is initialised with the uint32 (CHAR_MAX<<24)|(CHAR_MAX<<16)|(UCHAR_MAX<<8) (:205), stored little-endian: per pixel
//     (gradX, gradY, magnitude, pad) = (0, -1, 127, 127), so the min-|.| combination keeps gradX == 0 and |gradY| <= 1 at every scale.
// Like the reference object, the gradient and mask buffers persist between calls.
#pragma once
#include "lvo_common.hpp"

namespace lvo {

struct EdgeDetectorLBSP {
    int n_levels = 3;                 // EDGLBSP_DEFAULT_LEVEL_COUNT
    double hyst_low_factor = 0.5;     // EDGLBSP_DEFAULT_HYST_LOW_THRSH_FACT
    std::vector<uchar> grad, edge;    // m_vuLBSPGradMapData (4 bytes / px), m_vuEdgeTempMaskData (padded by 2 on every side)
    std::vector<std::vector<uchar>> pyr; // m_vvuInputPyrMaps: level l+1 image
    std::vector<std::pair<int, int>> sizes; // (rows, cols) per level

    static double default_threshold() { return 8.0 / 16.0; } // EDGLBSP_DEFAULT_DET_THRESHOLD = (MAX_GRAD_MAG/2)/MAX_GRAD_MAG

    /// apply_internal_lookup (:47-143): the pyramid. Level l+1 pixel (r/2, c/2), for even r and c, is the floor mean of the 16 LBSP
    /// neighbours of level-l pixel (r, c) per channel, or the pixel itself within the 2-px border. (The lookup maps themselves are
    /// recomputed where needed instead of being stored.)
    void build_pyramid(const uchar* img, int W, int H, int C) {
        if(!img || C < 1 || C > 4) throw std::runtime_error("input image must be non-empty and continuous, 8UC1 .. 8UC4"); // :144-160: 1 to 4 channels
        if(n_levels < 1) throw std::runtime_error("number of pyramid levels must be positive");
        sizes.assign(1, std::make_pair(H, W));
        for(int l = 1; l < n_levels; ++l) sizes.push_back(std::make_pair((sizes.back().first + 1) / 2, (sizes.back().second + 1) / 2));
        // the reference's loops run on size_t with `rows - 2` bounds: levels smaller than the LBSP patch are out of its domain
        if(sizes.back().first < 5 || sizes.back().second < 5) throw std::runtime_error("image too small for the number of pyramid levels");
        pyr.assign((size_t)std::max(n_levels - 1, 0), std::vector<uchar>());
        const uchar* cur = img;
        for(int l = 0; l + 1 < n_levels; ++l) {
            const int Hc = sizes[l].first, Wc = sizes[l].second, Hn = sizes[l + 1].first, Wn = sizes[l + 1].second;
            std::vector<uchar>& nxt = pyr[l];
            nxt.assign((size_t)Hn * Wn * C, 0);
            for(int r = 0; r < Hc; r += 2) for(int c = 0; c < Wc; c += 2) for(int k = 0; k < C; ++k) {
                uchar v;
                if(r < 2 || r >= Hc - 2 || c < 2 || c >= Wc - 2) v = cur[((size_t)r * Wc + c) * C + k];
                else {
                    uchar vals[16];
                    lbsp_lookup(cur, Wc, C, c, r, k, vals);
                    unsigned sum = 0;
                    for(int i = 0; i < 16; ++i) sum += vals[i];
                    v = (uchar)(sum / 16);
                }
                nxt[((size_t)(r / 2) * Wn + c / 2) * C + k] = v;
            }
            cur = nxt.data();
        }
    }

    /// apply_internal_threshold (:166-375) for one detection threshold in [0,16]
    void threshold_pass(const uchar* img, int W, int H, int C, uchar* out, uchar hi) {
        const uchar lo = (uchar)(hi * hyst_low_factor);
        const int mapW = W + 4, mapH = H + 4;
        const size_t grow = (size_t)mapW * 4, erow = (size_t)mapW;
        grad.resize((size_t)mapH * grow); edge.resize((size_t)mapH * erow); // new elements are zero, old ones keep their value
        std::fill(grad.begin(), grad.begin() + 2 * grow, 0);
        std::fill(grad.end() - 2 * grow, grad.end(), 0);
        std::fill(edge.begin(), edge.begin() + 2 * erow, 1);
        std::fill(edge.end() - 2 * erow, edge.end(), 1);
        for(size_t i = 2 * grow + 8; i + 4 <= (size_t)(mapH - 2) * grow - 8; i += 4) { grad[i] = 0x00; grad[i + 1] = 0xFF; grad[i + 2] = 0x7F; grad[i + 3] = 0x7F; }
        auto G = [&](int r, int c) -> uchar* { return grad.data() + (size_t)(r + 2) * grow + (size_t)(c + 2) * 4; }; // image coordinates
        auto E = [&](int r, int c) -> uchar* { return edge.data() + (size_t)(r + 2) * erow + (size_t)(c + 2); };     // padded row r+2
        auto minabs = [](signed char a, signed char b) { return std::abs((int)b) < std::abs((int)a) ? b : a; };       // std::min(a, b, |.|<|.|)
        std::vector<uchar*> stack;
        auto push = [&](uchar* p) { *p = 2; stack.push_back(p); };
        for(int l = n_levels - 1; l >= 0; --l) {
            const int Hc = sizes[l].first, Wc = sizes[l].second;
            const uchar* im = l ? pyr[l - 1].data() : img;
            for(int r = Hc - 1; r >= -2; --r) {
                if(r >= 0) {
                    for(int c = Wc - 1; c >= 0; --c) {
                        signed char gx = 0, gy = 0; uchar mag = 0;
                        if(!(r < 2 || r >= Hc - 2 || c < 2 || c >= Wc - 2)) lbsp_gradient_point(im, Wc, C, c, r, gx, gy, mag);
                        uchar* g = G(r, c);
                        g[0] = (uchar)minabs(gx, (signed char)g[0]);
                        g[1] = (uchar)minabs(gy, (signed char)g[1]);
                        g[2] = std::min(mag, g[2]);
                        if(l > 0) for(int dr = 0; dr < 2; ++dr) for(int dc = 0; dc < 2; ++dc) std::memcpy(G(2 * r + dr, 2 * c + dc), g, 4); // :258-269
                    }
                }
                if(l != 0) continue;
                std::memset(G(r, -2), 0, 8); std::memset(G(r, W), 0, 8);   // left / right padding of this gradient row (:273-274)
                if(r >= H - 2) continue;
                // non-maximum suppression of gradient row r+2, written to mask row r (sic)
                uchar* e = E(r, 0);
                e[-2] = e[-1] = 1; e[W] = e[W + 1] = 1;
                bool neighb_max = false;
                for(int c = 0; c < W; ++c) {
                    const uchar* g = G(r + 2, c);
                    const uchar mag = g[2];
                    auto M = [&](int dc, int dr) { return g[(ptrdiff_t)dc * 4 + (ptrdiff_t)dr * (ptrdiff_t)grow + 2]; };
                    auto horizontal = [&] { return mag > M(-1, 0) && mag > M(-2, 0) && mag >= M(1, 0) && mag >= M(2, 0); };
                    auto vertical = [&] { return mag > M(0, -1) && mag > M(0, -2) && mag >= M(0, 1) && mag >= M(0, 2); };
                    auto diagonal = [&](bool inv) { const int s = inv ? -1 : 1;
                        return mag > M(-s, -1) && mag > M(-2 * s, -2) && mag >= M(s, 1) && mag >= M(2 * s, 2); };
                    bool good = false;
                    if(mag >= lo) {
                        const signed char gx = (signed char)g[0], gy = (signed char)g[1];
                        const unsigned ax = (unsigned)std::abs((int)gx), ay = (unsigned)std::abs((int)gy) << 15;
                        const unsigned tg22 = ax * 13573u;            // tan(pi/8) in 1.15 fixed point (:300)
                        if(ay < tg22) good = horizontal();
                        else {
                            const unsigned tg67 = tg22 + (ax << 16);  // tan(3 pi/8) = tan(pi/8) + 2
                            if(ay > tg67) good = vertical();
                            else if(gx || gy) good = diagonal((((int)gx) ^ ((int)gy)) >= 0);
                            else good = diagonal(true) || diagonal(false);
                        }
                    }
                    if(!good) { neighb_max = false; e[c] = 1; continue; }
                    if(!neighb_max && mag >= hi && e[c + (ptrdiff_t)erow] != 2) { push(e + c); neighb_max = true; continue; }
                    e[c] = 0;
                }
            }
        }
        while(!stack.empty()) { // hysteresis (:353-372): 8-connected flood from the strong pixels through the "maybe" ones
            uchar* p = stack.back(); stack.pop_back();
            const ptrdiff_t rs = (ptrdiff_t)erow;
            const ptrdiff_t nb[8] = {-1, 1, -rs - 1, -rs, -rs + 1, rs - 1, rs, rs + 1};
            for(ptrdiff_t d : nb) {
                const ptrdiff_t q = (p - edge.data()) + d;
                // strong pixels of gradient rows 0 / 1 sit in padded mask rows 0 / 1 (the row shift): the reference then reads before its
                // buffer (undefined; a debug build asserts, :214-215). Here such reads count as "not an edge".
                if(q < 0 || q >= (ptrdiff_t)edge.size()) continue;
                if(!p[d]) push(p + d);
            }
        }
        for(int r = 0; r < H; ++r) for(int c = 0; c < W; ++c) out[(size_t)r * W + c] = (uchar)-(*E(r, c) >> 1);
    }

    /// apply_threshold (:391-410); thresholds outside [0,1] fall back to the default
    void apply_threshold(const uchar* img, int W, int H, int C, uchar* out, double thr) {
        if(thr < 0 || thr > 1) thr = default_threshold();
        build_pyramid(img, W, H, C);
        threshold_pass(img, W, H, C, out, (uchar)(thr * 16));
    }

    bool normalize_output = false;    // m_bNormalizeOutput (third constructor argument, default false)

    /// cv::normalize(x, x, 0, UCHAR_MAX, NORM_MINMAX) on an 8-bit map (:431-432; OpenCV core norm.cpp + convertTo): scale and shift in
    /// double, the conversion itself in float (multiply, add, round half to even, saturate). Pinned bit-exactly against cv2 4.13
    /// (tests/test_edge_oracle_cpu.py).
    static void normalize_minmax_u8(uchar* buf, size_t n) {
        if(!n) return;
        uchar lo = 255, hi = 0;
        for(size_t i = 0; i < n; ++i) { lo = std::min(lo, buf[i]); hi = std::max(hi, buf[i]); }
        const double smin = lo, smax = hi;
        const double scale = 255.0 * (smax - smin > 2.220446049250313e-16 ? 1. / (smax - smin) : 0.);
        const double shift = 0.0 - smin * scale;
        const float a = (float)scale, b = (float)shift;
        for(size_t i = 0; i < n; ++i) {
            volatile float m = (float)buf[i] * a;   // two roundings, as OpenCV's scalar and SIMD paths produce here (checked against cv2)
            const float v = m + b;
            const long r = std::lrint(v);
            buf[i] = (uchar)(r < 0 ? 0 : r > 255 ? 255 : r);
        }
    }

    /// apply (:412-433): every threshold 0..15 contributes saturate(cvRound(255 / 16.0)) = 16 where it finds an edge; optional normalisation
    void apply(const uchar* img, int W, int H, int C, uchar* out) {
        build_pyramid(img, W, H, C);
        std::vector<uchar> tmp((size_t)W * H);
        std::memset(out, 0, (size_t)W * H);
        for(int t = 0; t < 16; ++t) {
            threshold_pass(img, W, H, C, tmp.data(), (uchar)t);
            for(size_t i = 0; i < tmp.size(); ++i) out[i] = (uchar)std::min(255, (int)out[i] + (tmp[i] ? 16 : 0));
        }
        if(normalize_output) normalize_minmax_u8(out, (size_t)W * H);
    }
};

} // namespace lvo
