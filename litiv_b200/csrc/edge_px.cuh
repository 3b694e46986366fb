// litiv_b200 — per-pixel bodies of the LBSP edge detector (reference imgproc/src/EdgeDetectorLBSP.cpp:47-375), written as plain
// host/device functions so that the SAME code runs inside the kernels of edge.cuh and inside the CPU emulation the "not gpu" tests
// compile with g++ (tests/edge_emul.cpp) to check the parallel restatement against the sequential CPU restatement of the tests.
//
// The reference walks one shared gradient map bottom-up, coarse level first, each level min-combining its own gradient with what the
// coarser level replicated 2x2 into the same map (:243-269), then suppresses non-maxima row by row into a padded mask and floods the
// "maybe" pixels from the strong ones with a stack (:276-372). Order-free form used here:
//   V_l(r,c)   = combine(grad_l(r,c), l == L-1 ? (0,-1,127) : V_{l+1}(r/2, c/2))         one uchar4 map per level
//   Mag(r,c)   = V_0(r,c).z inside the image; V_1(H1-1, c/2).z on row H when H is odd (the replicated coarse row the reference
//                leaves in its bottom padding and never clears: read by the 5x5 suppression of rows H-2, H-1); 0 elsewhere
//   class      = 1 (no edge) / 0 (maybe) / 2 (strong), gradient row r+2 classified into mask row r (the reference's row shift, :263-272)
//   mask rows H-2, H-1 are never written by the suppression: they persist from call to call (values 0 / 1 / 3 = a 2 of an earlier call)
//   seeds      = strong pixels, minus those of mask row H-3 that sit above a persisted 2 (:333: `below != 2`); the demotions the
//                reference applies elsewhere (`neighb_max`, a fresh 2 below) always leave the pixel 8-adjacent to a seed, so they do
//                not change the flooded set
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#ifndef LVB_HD
#define LVB_HD __host__ __device__ __forceinline__
#endif

namespace lvb_edge {

typedef unsigned char uchar;
enum : uchar { EDGE_MAYBE = 0, EDGE_NONE = 1, EDGE_YES = 2, EDGE_STALE_YES = 3 };

/// apply_internal_lookup (:84-124): pixel (r,c) of level l (r, c even) -> pixel (r/2, c/2) of level l+1: the pixel itself within the
/// 2-px border, else the floor mean of its 16 LBSP neighbours (LBSP.hpp:275-298: the 5x5 double cross)
LVB_HD uchar pyr_down_px(const uchar* cur, size_t pitch, int Wc, int Hc, int C, int r, int c, int k) {
    const uchar* p = cur + (size_t)r * pitch + (size_t)c * C + k;
    if(r < 2 || r >= Hc - 2 || c < 2 || c >= Wc - 2) return *p;
    const ptrdiff_t P = (ptrdiff_t)pitch;
    unsigned s = 0;
    s += p[-2 * C] + p[2 * C] + p[-2 * P] + p[2 * P];                                   // (+-2, 0), (0, +-2)
    s += p[-2 * P - 2 * C] + p[-2 * P + 2 * C] + p[2 * P - 2 * C] + p[2 * P + 2 * C];   // (+-2, +-2)
    s += p[-C] + p[C] + p[-P] + p[P];                                                   // (+-1, 0), (0, +-1)
    s += p[-P - C] + p[-P + C] + p[P - C] + p[P + C];                                   // (+-1, +-1)
    return (uchar)(s >> 4);
}

/// :255-257 with USE_MIN_GRAD_ORIENT: std::min(new, old, |.| < |.|) keeps the NEW value on ties; the magnitude is a plain min
LVB_HD uchar4 edge_combine(uchar4 own, uchar4 coarse) {
    const int nx = (signed char)own.x, ox = (signed char)coarse.x, ny = (signed char)own.y, oy = (signed char)coarse.y;
    uchar4 o;
    o.x = (uchar)(signed char)((ox < 0 ? -ox : ox) < (nx < 0 ? -nx : nx) ? ox : nx);
    o.y = (uchar)(signed char)((oy < 0 ? -oy : oy) < (ny < 0 ? -ny : ny) ? oy : ny);
    o.z = own.z < coarse.z ? own.z : coarse.z;
    o.w = coarse.w;   // the pad byte is never written by the level loop (:255-257 touch three bytes): it keeps the initial 127
    return o;
}
/// the little-endian bytes of (CHAR_MAX<<24)|(CHAR_MAX<<16)|(UCHAR_MAX<<8) (:205)
LVB_HD uchar4 edge_init_value() { uchar4 o; o.x = 0; o.y = 0xFF; o.z = 0x7F; o.w = 0x7F; return o; }

struct EdgeMaps { const uchar4* V0; int W, H; const uchar4* V1; int W1, H1; }; // V1 == null when the detector has one level

/// gradient magnitude as the suppression loop sees it at image coordinates (r,c), r in [-2, H+1], c in [-2, W+1]
LVB_HD unsigned edge_mag(const EdgeMaps& m, int r, int c) {
    if(r < 0 || c < 0) return 0;
    if(r < m.H) return c < m.W ? m.V0[(size_t)r * m.W + c].z : 0u;
    if(r == m.H && (m.H & 1) && m.V1 && c < 2 * m.W1) return m.V1[(size_t)(m.H1 - 1) * m.W1 + (c >> 1)].z;
    return 0;
}

/// non-maximum suppression of gradient pixel (rg, c) (:282-349; USE_5x5_NON_MAX_SUPP, USE_3_AXIS_ORIENT): EDGE_NONE / EDGE_MAYBE / EDGE_YES
LVB_HD uchar edge_nms_class(const EdgeMaps& m, int rg, int c, unsigned lo, unsigned hi) {
    const uchar4 g = m.V0[(size_t)rg * m.W + c];
    const unsigned mag = g.z;
    if(mag < lo) return EDGE_NONE;
    const int gx = (signed char)g.x, gy = (signed char)g.y;
    const unsigned ax = (unsigned)(gx < 0 ? -gx : gx), ay = (unsigned)(gy < 0 ? -gy : gy) << 15;
    const unsigned tg22 = ax * 13573u;
    bool good;
#define LVB_M(dc, dr) edge_mag(m, rg + (dr), c + (dc))
    if(ay < tg22) good = mag > LVB_M(-1, 0) && mag > LVB_M(-2, 0) && mag >= LVB_M(1, 0) && mag >= LVB_M(2, 0);
    else if(ay > tg22 + (ax << 16)) good = mag > LVB_M(0, -1) && mag > LVB_M(0, -2) && mag >= LVB_M(0, 1) && mag >= LVB_M(0, 2);
    else {
        const bool d_inv = mag > LVB_M(1, -1) && mag > LVB_M(2, -2) && mag >= LVB_M(-1, 1) && mag >= LVB_M(-2, 2);   // s = -1
        const bool d_std = mag > LVB_M(-1, -1) && mag > LVB_M(-2, -2) && mag >= LVB_M(1, 1) && mag >= LVB_M(2, 2);   // s = +1
        if(gx || gy) good = ((gx ^ gy) >= 0) ? d_inv : d_std;
        else good = d_inv || d_std;
    }
#undef LVB_M
    if(!good) return EDGE_NONE;
    return mag >= hi ? EDGE_YES : EDGE_MAYBE;
}

/// the value mask row r (r <= H-3) gets before the flood: the class of gradient row r+2, a strong pixel of row H-3 demoted to "maybe"
/// when the persisted row below holds a 2 of an earlier call (stored as EDGE_STALE_YES)
LVB_HD uchar edge_mask_value(const EdgeMaps& m, int r, int c, unsigned lo, unsigned hi, const uchar* mask) {
    uchar k = edge_nms_class(m, r + 2, c, lo, hi);
    if(k == EDGE_YES && r == m.H - 3 && mask[(size_t)(m.H - 2) * m.W + c] == EDGE_STALE_YES) k = EDGE_MAYBE;
    return k;
}

} // namespace lvb_edge
