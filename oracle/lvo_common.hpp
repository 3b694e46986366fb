// ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the shipped product path.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
//
// CPU restatement of the LITIV change-detection hot path (plstcharles/litiv), written from the
// reference's behaviour, not copied from it.  Every function cites the reference file:line it follows
// (paths relative to the reference root, modules/...).
//
// Parity status: PINNED. LBSP by the reference's own golden vector (features2d/test/data/test_lbsp.bin, see
// tests/golden/); SuBSENSE / LOBSTER / PAWCS / ViBe / PBAS (initialize, apply, refreshModel, getBackgroundImage) and EdgeDetectorLBSP by
// the reference itself: oracle/_ref/liblitiv_ref.so is built from the reference's own unmodified sources against
// oracle/cvcompat (`make _ref`), and tests/test_ref_pin_cpu.py holds the reference-order mode of these
// restatements equal to it bit for bit (masks, models, float maps). The OpenCV-equivalent image operations are
// cross-checked against cv2 in tests/.
#pragma once
#include <cstdint>
#include <cstddef>
#include <cstring>
#include <cmath>
#include <vector>
#include <string>
#include <algorithm>
#include <stdexcept>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

namespace lvo {

typedef unsigned char uchar;
typedef unsigned short ushort;

// ---------------------------------------------------------------------------------------------
// RNGs
// ---------------------------------------------------------------------------------------------

/// glibc rand()/srand() clone (TYPE_3 additive feedback, degree 31, separation 3).  The reference
/// calls libc rand() directly (18 call sites in SuBSENSE.cpp, e.g. video/src/BackgroundSubtractorSuBSENSE.cpp:86,93,264).
struct GlibcRand {
    int f, b; // front/rear indices into the 31-entry ring
    int32_t ring[31];
    explicit GlibcRand(unsigned seed = 1) { srand(seed); }
    void srand(unsigned seed) {
        if(seed == 0) seed = 1;
        ring[0] = (int32_t)seed;
        for(int i = 1; i < 31; ++i) {
            // 16807 * prev mod (2^31-1), Schrage's method as glibc does it
            long hi = ring[i-1] / 127773;
            long lo = ring[i-1] % 127773;
            long word = 16807 * lo - 2836 * hi;
            if(word < 0) word += 2147483647;
            ring[i] = (int32_t)word;
        }
        f = 3; b = 0;
        for(int i = 0; i < 310; ++i) (void)next();
    }
    int next() {
        uint32_t v = (uint32_t)ring[f] + (uint32_t)ring[b];
        ring[f] = (int32_t)v;
        int result = (int)(v >> 1);
        if(++f >= 31) f = 0;
        if(++b >= 31) b = 0;
        return result;
    }
};

/// Philox4x32-10 (Salmon et al., SC'11), the counter-based generator the snapshot mode and the GPU use.
inline void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for(int i = 0; i < 10; ++i) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        const uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/// RNG draw domains (counter word 3) shared with the CUDA kernels (litiv_b200/csrc/philox.cuh)
enum { DOM_APPLY = 0, DOM_REFRESH = 1, DOM_REFRESH_START = 2, DOM_PAWCS_A = 3, DOM_PAWCS_B = 4 };

/// 31-bit draw #site for (frame,pixel) — same range as libc rand()
inline int philox_draw(uint64_t seed, uint32_t frame, uint32_t pixel, uint32_t site, uint32_t domain) {
    const uint32_t ctr[4] = {frame, pixel, site >> 2, domain};
    const uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t out[4];
    philox4x32_10(ctr, key, out);
    return (int)(out[site & 3] >> 1);
}

// ---------------------------------------------------------------------------------------------
// small math (utils/include/litiv/utils/math.hpp)
// ---------------------------------------------------------------------------------------------

/// math.hpp:199-203 (integer L1dist; returns the unsigned type of the input)
inline uchar L1dist_u8(uchar a, uchar b) { return (uchar)std::abs((int)a - (int)b); }
/// math.hpp:228-234: array overload accumulates in and returns uchar => wraps mod 256 (quirk Q1)
template<int C> inline uchar L1dist_arr_u8(const uchar* a, const uchar* b) {
    uchar r = 0;
    for(int c = 0; c < C; ++c) r = (uchar)(r + L1dist_u8(a[c], b[c]));
    return r;
}
/// math.hpp:627-700
inline int popcount16(ushort x) { return __builtin_popcount((unsigned)x); }
/// math.hpp:702-735
inline int hdist16(ushort a, ushort b) { return __builtin_popcount((unsigned)(a ^ b)); }

/// math.hpp:474-496 colour distortion (integer)
template<int C> inline size_t cdist_u8(const uchar* curr, const uchar* bg) {
    bool nonconst = false;
    bool nonnull = (curr[0] != bg[0]);
    for(int c = 1; c < C; ++c) {
        nonconst |= (curr[c] != curr[c-1]) || (bg[c] != bg[c-1]);
        nonnull |= (curr[c] != bg[c]);
    }
    if(!nonconst || !nonnull) return 0;
    uint64_t curr_sqr = 0, bg_sqr = 0, mix = 0;
    for(int c = 0; c < C; ++c) {
        curr_sqr += (uint64_t)(curr[c] * curr[c]);
        bg_sqr += (uint64_t)(bg[c] * bg[c]);
        mix += (uint64_t)(curr[c] * bg[c]);
    }
    const float d = (float)(curr_sqr - (mix * mix) / std::max(bg_sqr, (uint64_t)1));
    return (size_t)std::sqrt(d);
}
/// math.hpp:596-605 ; with the pointer/array overloads the L1 term is the wrapping uchar one (Q1)
inline size_t cmixdist(size_t l1, size_t cd) { return l1 / 2 + cd * 4; }

/// cv::saturate_cast<uchar>(float): cvRound (round-half-even) then clamp
inline uchar sat_u8(float v) {
    const long r = std::lrint((double)v); // default rounding mode: to nearest even
    return (uchar)(r < 0 ? 0 : (r > 255 ? 255 : r));
}
inline uchar sat_u8_int(long v) { return (uchar)(v < 0 ? 0 : (v > 255 ? 255 : v)); }

// ---------------------------------------------------------------------------------------------
// LBSP (features2d/include/litiv/features2d/LBSP.hpp)
// ---------------------------------------------------------------------------------------------

/// LBSP.hpp:292-294: bit n of the descriptor <-> neighbour offset (dx[n],dy[n])
static const int LBSP_DX[16] = {-2, 2, 0, 0, -2, 2, 2, -2, 0, -1, 0, 1, -1, 1, 1, -1};
static const int LBSP_DY[16] = { 0, 0,-2, 2,  2,-2, 2, -2, 1,  0,-1, 0, -1, 1,-1,  1};

/// LBSP.hpp:300-319 (lookup_16bits_dbcross): gather 16 neighbours of channel c
inline void lbsp_lookup(const uchar* img, int W, int C, int x, int y, int c, uchar vals[16]) {
    const size_t rs = (size_t)W * C;
    const uchar* p = img + (size_t)y * rs + (size_t)x * C + c;
    for(int n = 0; n < 16; ++n)
        vals[n] = p[(ptrdiff_t)rs * LBSP_DY[n] + (ptrdiff_t)C * LBSP_DX[n]];
}
/// LBSP.hpp:193-224 (computeDescriptor_threshold): strict '>' on the absolute difference.
/// Scalar definition plus an SSE2 form (the reference also uses SSE for this step, LBSP.hpp:203-223), so the
/// CPU baseline timing is not handicapped; both give identical bits (checked in tests/test_oracle_helpers.py).
inline ushort lbsp_threshold_scalar(const uchar vals[16], uchar ref, uchar t) {
    unsigned d = 0;
    for(int n = 0; n < 16; ++n)
        d |= (unsigned)(L1dist_u8(vals[n], ref) > t) << n;
    return (ushort)d;
}
inline ushort lbsp_threshold(const uchar vals[16], uchar ref, uchar t) {
#if defined(__SSE2__)
    const __m128i v = _mm_loadu_si128((const __m128i*)vals);
    const __m128i r = _mm_set1_epi8((char)ref);
    const __m128i d = _mm_or_si128(_mm_subs_epu8(v, r), _mm_subs_epu8(r, v));         // |v-ref|
    const __m128i over = _mm_subs_epu8(d, _mm_set1_epi8((char)t));                       // >0 iff d>t
    return (ushort)(~_mm_movemask_epi8(_mm_cmpeq_epi8(over, _mm_setzero_si128())) & 0xFFFF);
#else
    return lbsp_threshold_scalar(vals, ref, t);
#endif
}

/// LBSP::computeDescriptor_gradient<nChannels, nAbsOffset=20, nRelShift=2> (features2d/include/litiv/features2d/LBSP.hpp:235-256):
/// per channel the pattern is thresholded with t = ((ref >> 2) + 20) / 2; the channel with the largest popcount wins (the scan starts
/// from the LAST channel and replaces on strictly greater); gradX / gradY = popcount differences over the masks of LBSP.hpp:288-291.
static const ushort LBSP_GRADX_POS = (1<<0)|(1<<4)|(1<<7)|(1<<9)|(1<<12)|(1<<15), LBSP_GRADX_NEG = (1<<1)|(1<<5)|(1<<6)|(1<<11)|(1<<13)|(1<<14);
static const ushort LBSP_GRADY_POS = (1<<3)|(1<<4)|(1<<6)|(1<<8)|(1<<13)|(1<<15), LBSP_GRADY_NEG = (1<<2)|(1<<5)|(1<<7)|(1<<10)|(1<<12)|(1<<14);
inline void lbsp_gradient_point(const uchar* img, int W, int C, int x, int y, signed char& gx, signed char& gy, uchar& mag) {
    ushort best = 0; int best_mag = -1;
    for(int k = 0; k < C; ++k) {
        const int c = k == 0 ? C - 1 : k - 1; // channel order of the reference: last, then 0 .. C-2
        uchar vals[16];
        lbsp_lookup(img, W, C, x, y, c, vals);
        const uchar ref = img[((size_t)y * W + x) * C + c];
        const ushort d = lbsp_threshold(vals, ref, (uchar)((((int)ref >> 2) + 20) / 2));
        const int m = __builtin_popcount((unsigned)d);
        if(best_mag < m) { best_mag = m; best = d; }
    }
    mag = (uchar)best_mag;
    gx = (signed char)(__builtin_popcount((unsigned)(best & LBSP_GRADX_POS)) - __builtin_popcount((unsigned)(best & LBSP_GRADX_NEG)));
    gy = (signed char)(__builtin_popcount((unsigned)(best & LBSP_GRADY_POS)) - __builtin_popcount((unsigned)(best & LBSP_GRADY_NEG)));
}
/// dense map in the layout of EdgeDetectorLBSP's gradient map (imgproc/src/EdgeDetectorLBSP.cpp:196: 4 bytes per pixel = gradX, gradY,
/// magnitude, 0); pixels within the 2-px border have the all-equal lookup of :84-100, i.e. a zero pattern: (0, 0, 0, 0)
inline void lbsp_gradient_dense(const uchar* img, int W, int H, int C, uchar* out) {
    std::memset(out, 0, (size_t)W * H * 4);
    for(int y = 2; y < H - 2; ++y) for(int x = 2; x < W - 2; ++x) {
        signed char gx, gy; uchar mag;
        lbsp_gradient_point(img, W, C, x, y, gx, gy, mag);
        uchar* o = out + ((size_t)y * W + x) * 4;
        o[0] = (uchar)gx; o[1] = (uchar)gy; o[2] = mag;
    }
}

/// The 2-px border of the output is left untouched (the reference never writes it); we keep whatever
/// the caller put there (tests pre-fill zeros).
/// features2d/src/LBSP.cpp:102-152 (lbsp_computeImpl, dense; abs or rel threshold; optional ref image)
inline void lbsp_compute_dense(const uchar* img, const uchar* ref_or_null, int W, int H, int C,
                               bool use_rel, float rel, size_t thr, ushort* out) {
    const uchar* ref = ref_or_null ? ref_or_null : img;
    const uchar tabs = sat_u8_int((long)(int)thr);
    for(int y = 2; y < H - 2; ++y) {
        for(int x = 2; x < W - 2; ++x) {
            for(int c = 0; c < C; ++c) {
                const uchar r = ref[((size_t)y * W + x) * C + c];
                const uchar t = use_rel ? sat_u8((float)r * rel + (float)thr) : tabs;
                uchar vals[16];
                lbsp_lookup(img, W, C, x, y, c, vals);
                out[((size_t)y * W + x) * C + c] = lbsp_threshold(vals, r, t);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// sampling patterns (utils/include/litiv/utils/opencv.hpp)
// ---------------------------------------------------------------------------------------------

/// opencv.hpp:859-870
inline void clamp_coords(int& x, int& y, int border, int W, int H) {
    if(x < border) x = border; else if(x >= W - border) x = W - border - 1;
    if(y < border) y = border; else if(y >= H - border) y = H - border - 1;
}
/// opencv.hpp:909-925 table, walk of :873-891
static const int PATTERN_7x7[7][7] = {
    {2, 4, 6, 7, 6, 4, 2}, {4, 8, 12, 14, 12, 8, 4}, {6, 12, 21, 25, 21, 12, 6}, {7, 14, 25, 28, 25, 14, 7},
    {6, 12, 21, 25, 21, 12, 6}, {4, 8, 12, 14, 12, 8, 4}, {2, 4, 6, 7, 6, 4, 2}};
inline void sample_pos_7x7(int rnd, int& sx, int& sy, int ox, int oy, int border, int W, int H) {
    int r = 1 + (rnd % 512);
    for(sy = 0; sy < 7; ++sy) {
        for(sx = 0; sx < 7; ++sx) {
            r -= PATTERN_7x7[sy][sx];
            if(r <= 0) goto stop;
        }
    }
stop:
    sx += ox - 3; sy += oy - 3;
    clamp_coords(sx, sy, border, W, H);
}
/// opencv.hpp:941-951
static const int NB3[8][2] = {{-1, 1}, {0, 1}, {1, 1}, {-1, 0}, {1, 0}, {-1, -1}, {0, -1}, {1, -1}};
/// opencv.hpp:954-966
static const int NB5[24][2] = {
    {-2, 2}, {-1, 2}, {0, 2}, {1, 2}, {2, 2}, {-2, 1}, {-1, 1}, {0, 1}, {1, 1}, {2, 1}, {-2, 0}, {-1, 0},
    {1, 0}, {2, 0}, {-2, -1}, {-1, -1}, {0, -1}, {1, -1}, {2, -1}, {-2, -2}, {-1, -2}, {0, -2}, {1, -2}, {2, -2}};
inline void neighbor_pos_3x3(int rnd, int& nx, int& ny, int ox, int oy, int border, int W, int H) {
    const int r = rnd % 8; nx = ox + NB3[r][0]; ny = oy + NB3[r][1]; clamp_coords(nx, ny, border, W, H);
}
inline void neighbor_pos_5x5(int rnd, int& nx, int& ny, int ox, int oy, int border, int W, int H) {
    const int r = rnd % 24; nx = ox + NB5[r][0]; ny = oy + NB5[r][1]; clamp_coords(nx, ny, border, W, H);
}

// ---------------------------------------------------------------------------------------------
// OpenCV-equivalent mask ops on byte images (semantics: SURVEY.md Appendix E, cross-checked vs cv2 in tests)
// ---------------------------------------------------------------------------------------------

/// cv::dilate / cv::erode with a (2r+1)x(2r+1) rect (== `iterations=r` of the default 3x3); pixels outside
/// the image are ignored (default morphology border value). Separable implementation.
inline void morph_rect(const uchar* src, uchar* dst, int W, int H, int r, bool dilate) {
    std::vector<uchar> tmp((size_t)W * H);
    for(int y = 0; y < H; ++y) {
        const uchar* s = src + (size_t)y * W; uchar* t = tmp.data() + (size_t)y * W;
        for(int x = 0; x < W; ++x) {
            const int x0 = std::max(0, x - r), x1 = std::min(W - 1, x + r);
            uchar v = s[x0];
            if(dilate) { for(int k = x0 + 1; k <= x1; ++k) v = std::max(v, s[k]); }
            else       { for(int k = x0 + 1; k <= x1; ++k) v = std::min(v, s[k]); }
            t[x] = v;
        }
    }
    for(int y = 0; y < H; ++y) {
        const int y0 = std::max(0, y - r), y1 = std::min(H - 1, y + r);
        uchar* d = dst + (size_t)y * W;
        std::memcpy(d, tmp.data() + (size_t)y0 * W, (size_t)W);
        for(int k = y0 + 1; k <= y1; ++k) {
            const uchar* t = tmp.data() + (size_t)k * W;
            if(dilate) { for(int x = 0; x < W; ++x) d[x] = std::max(d[x], t[x]); }
            else       { for(int x = 0; x < W; ++x) d[x] = std::min(d[x], t[x]); }
        }
    }
}

/// cv::medianBlur on a {0,255} mask == majority vote over the kxk window with replicated borders
inline void median_binary(const uchar* src, uchar* dst, int W, int H, int k) {
    const int r = k / 2; const int half = (k * k) / 2;
    std::vector<ushort> hs((size_t)W * H);
    for(int y = 0; y < H; ++y) {
        const uchar* s = src + (size_t)y * W; ushort* h = hs.data() + (size_t)y * W;
        int acc = 0;
        for(int d = -r; d <= r; ++d) acc += s[std::min(std::max(d, 0), W - 1)] ? 1 : 0;
        h[0] = (ushort)acc;
        for(int x = 1; x < W; ++x) {
            acc += (s[std::min(x + r, W - 1)] ? 1 : 0) - (s[std::max(x - r - 1, 0)] ? 1 : 0);
            h[x] = (ushort)acc;
        }
    }
    std::vector<int> col((size_t)W, 0);
    for(int d = -r; d <= r; ++d) {
        const ushort* h = hs.data() + (size_t)std::min(std::max(d, 0), H - 1) * W;
        for(int x = 0; x < W; ++x) col[x] += h[x];
    }
    for(int y = 0; y < H; ++y) {
        if(y > 0) {
            const ushort* ha = hs.data() + (size_t)std::min(y + r, H - 1) * W;
            const ushort* hb = hs.data() + (size_t)std::max(y - r - 1, 0) * W;
            for(int x = 0; x < W; ++x) col[x] += (int)ha[x] - (int)hb[x];
        }
        uchar* d = dst + (size_t)y * W;
        for(int x = 0; x < W; ++x) d[x] = col[x] > half ? 255 : 0;
    }
}

/// cv::floodFill(img, Point(0,0), 255) with default 4-connectivity and zero lo/up diffs on a {0,255} mask:
/// every pixel 4-connected to (0,0) through pixels equal to img(0,0) becomes 255.
inline void floodfill_from_origin(uchar* img, int W, int H) {
    const uchar v0 = img[0];
    if(v0 == 255) return;
    std::vector<int> stack;
    stack.push_back(0); img[0] = 255;
    while(!stack.empty()) {
        const int p = stack.back(); stack.pop_back();
        const int x = p % W, y = p / W;
        if(x > 0 && img[p - 1] == v0) { img[p - 1] = 255; stack.push_back(p - 1); }
        if(x < W - 1 && img[p + 1] == v0) { img[p + 1] = 255; stack.push_back(p + 1); }
        if(y > 0 && img[p - W] == v0) { img[p - W] = 255; stack.push_back(p - W); }
        if(y < H - 1 && img[p + W] == v0) { img[p + W] = 255; stack.push_back(p + W); }
    }
}

/// cv::resize(INTER_AREA) u8 -> u8 for an exact integer shrink factor s: rint_half_even(sum/(s*s))
inline void resize_area_exact(const uchar* src, int W, int H, int C, int s, uchar* dst) {
    const int w = W / s, h = H / s;
    const float scale = 1.0f / (float)(s * s);
    for(int y = 0; y < h; ++y) for(int x = 0; x < w; ++x) for(int c = 0; c < C; ++c) {
        int sum = 0;
        for(int dy = 0; dy < s; ++dy) for(int dx = 0; dx < s; ++dx)
            sum += src[((size_t)(y * s + dy) * W + (x * s + dx)) * C + c];
        dst[((size_t)y * w + x) * C + c] = sat_u8((float)sum * scale);
    }
}

/// cv::resize(INTER_AREA) u8 -> u8 for a non-integer shrink factor: OpenCV's general area path (imgproc resize.cpp:
/// computeResizeAreaTab + ResizeArea_Invoker<uchar,float>), restated with its float accumulation order: per source row,
/// buf[dx] += S[sx]*alpha over the column table; per destination row, sum = beta0*buf0 then += beta*buf; round-half-even +
/// saturate at the end. OpenCV is not vendored by the reference; this restatement is pinned against cv2 4.13 in
/// tests/test_oracle_cpu.py (bit-exact on CDnet's non-multiple-of-8 frame sizes).
struct AreaTabEntry { int di, si; float alpha; };
inline std::vector<AreaTabEntry> area_tab(int ssize, int dsize, double scale) {
    std::vector<AreaTabEntry> tab;
    for(int dx = 0; dx < dsize; ++dx) {
        const double fsx1 = dx * scale, fsx2 = fsx1 + scale, cell = std::min(scale, ssize - fsx1);
        int sx1 = (int)std::ceil(fsx1), sx2 = (int)std::floor(fsx2);
        sx2 = std::min(sx2, ssize - 1); sx1 = std::min(sx1, sx2);
        if(sx1 - fsx1 > 1e-3) tab.push_back({dx, sx1 - 1, (float)((sx1 - fsx1) / cell)});
        for(int sx = sx1; sx < sx2; ++sx) tab.push_back({dx, sx, (float)(1.0 / cell)});
        if(fsx2 - sx2 > 1e-3) tab.push_back({dx, sx2, (float)(std::min(std::min(fsx2 - sx2, 1.), cell) / cell)});
    }
    return tab;
}
inline void resize_area_general(const uchar* src, int W, int H, int C, int dw, int dh, uchar* dst) {
    const double scale_x = 1. / ((double)dw / W), scale_y = 1. / ((double)dh / H);
    const std::vector<AreaTabEntry> xt = area_tab(W, dw, scale_x), yt = area_tab(H, dh, scale_y);
    std::vector<float> buf((size_t)dw * C), sum((size_t)dw * C, 0.f);
    int prev = -1;
    auto flush = [&](int dy) { for(int i = 0; i < dw * C; ++i) dst[(size_t)dy * dw * C + i] = sat_u8(sum[i]); };
    for(const AreaTabEntry& ye : yt) {
        std::fill(buf.begin(), buf.end(), 0.f);
        const uchar* S = src + (size_t)ye.si * W * C;
        for(const AreaTabEntry& xe : xt)
            for(int c = 0; c < C; ++c) buf[(size_t)xe.di * C + c] += (float)S[(size_t)xe.si * C + c] * xe.alpha;
        if(ye.di != prev) {
            if(prev >= 0) flush(prev);
            for(int i = 0; i < dw * C; ++i) sum[i] = ye.alpha * buf[i];
            prev = ye.di;
        } else for(int i = 0; i < dw * C; ++i) sum[i] += ye.alpha * buf[i];
    }
    if(prev >= 0) flush(prev);
}

/// IIBackgroundSubtractor::initialize_common ROI handling (video/src/BackgroundSubtractionUtils.cpp:82-99)
/// and validateROI (:28-36). Returns the final ROI ({0,128,255}); orig_count = countNonZero before validateROI.
inline void build_roi(const uchar* roi_or_null, int W, int H, int border, std::vector<uchar>& roi, size_t& orig_count, size_t& final_count) {
    roi.assign((size_t)W * H, 255);
    if(roi_or_null) {
        for(size_t i = 0; i < (size_t)W * H; ++i)
            if(roi_or_null[i] != 0 && roi_or_null[i] != 255) throw std::runtime_error("provided ROI mat values must be 0 or 255 only");
        std::vector<uchar> dil((size_t)W * H);
        morph_rect(roi_or_null, dil.data(), W, H, border, true);
        for(size_t i = 0; i < (size_t)W * H; ++i) roi[i] = roi_or_null[i] | (dil[i] ? 128 : 0); // 255/2 saturate-rounds to 128
    }
    orig_count = 0;
    for(size_t i = 0; i < (size_t)W * H; ++i) orig_count += roi[i] != 0;
    if(orig_count == 0) throw std::runtime_error("provided ROI mat contains no useful pixels");
    for(int y = 0; y < H; ++y) for(int x = 0; x < W; ++x)
        if(x < border || y < border || x >= W - border || y >= H - border) roi[(size_t)y * W + x] = 0;
    final_count = 0;
    for(size_t i = 0; i < (size_t)W * H; ++i) final_count += roi[i] != 0;
    if(final_count == 0) throw std::runtime_error("provided ROI mat contains no useful pixels away from borders");
}

/// IBackgroundSubtractorLBSP_::initialize_common LUT (video/src/BackgroundSubtractorLBSP.cpp:29-30, 42-43; quirk Q2)
inline void build_lbsp_lut(int C, float rel, size_t off, uchar lut[256]) {
    for(size_t t = 0; t < 256; ++t)
        lut[t] = (C == 1) ? sat_u8(((float)t * rel + (float)off) / 3) : sat_u8((float)t * rel + (float)off);
}

enum Mode { MODE_REFERENCE = 0, MODE_SNAPSHOT = 1 };

struct NamedBuf { const char* name; void* ptr; size_t bytes; };

} // namespace lvo
