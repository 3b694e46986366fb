// litiv_b200 — CDnet-style binary classification counters on the device: replaces
// lv::BinClassif::accumulate(oClassif, oGT, oROI) (reference modules/datasets/src/metrics.cpp:21-61; label values
// datasets/include/litiv/datasets/metrics.hpp:23-27). This is the step right after apply() in the reference's evaluation
// loop; done on the device it reads the instance's bit-packed final mask, so scoring a sequence needs no mask read-back.
#pragma once
#include "state.cuh"

namespace lvb {

enum { BC_TP = 0, BC_TN = 1, BC_FP = 2, BC_FN = 3, BC_SE = 4, BC_DC = 5, BC_COUNT = 6 }; // BinClassif::CountersList (metrics.hpp:40-48)
constexpr uint32_t GT_POSITIVE = 255u, GT_NEGATIVE = 0u, GT_OUTOFSCOPE = 85u, GT_UNKNOWN = 170u, GT_SHADOW = 50u;

struct BinClassifArgs {
    int W, H, WW;
    const uchar* classif; size_t cpitch;   // byte mask (positive == 255), or null: use classif_bits
    const uint32_t* classif_bits;          // bit-packed mask [H][WW] (bit set == 255)
    const uchar* gt; size_t gpitch;        // CDnet labels
    const uchar* roi; size_t rpitch;       // evaluation ROI or null (pixels equal to 0 are don't-care)
    unsigned long long* counters;          // [BC_COUNT], added to
};

/// one warp per 32-pixel mask word and row: 6 ballots give the warp's counters, lane 0 of each warp adds them to shared memory,
/// one thread per CTA to global memory
__global__ void __launch_bounds__(256) binclassif_kernel(const BinClassifArgs A) {
    __shared__ unsigned int s_cnt[BC_COUNT];
    const int tid = threadIdx.y * 32 + threadIdx.x;
    if(tid < BC_COUNT) s_cnt[tid] = 0;
    __syncthreads();
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    const bool in = x < A.W && y < A.H;
    bool tp = false, tn = false, fp = false, fn = false, se = false, dc = false;
    if(in) {
        const uint32_t g = A.gt[(size_t)y * A.gpitch + x];
        const bool pos = A.classif ? (A.classif[(size_t)y * A.cpitch + x] == 255u) : (((A.classif_bits[(size_t)y * A.WW + (x >> 5)] >> (x & 31)) & 1u) != 0u);
        const bool scored = g != GT_OUTOFSCOPE && g != GT_UNKNOWN && (!A.roi || A.roi[(size_t)y * A.rpitch + x] != GT_NEGATIVE);
        if(scored) {
            const bool gpos = g == GT_POSITIVE;
            tp = pos && gpos; fp = pos && !gpos; fn = !pos && gpos; tn = !pos && !gpos;
            se = pos && g == GT_SHADOW;
        } else dc = true;
    }
    const uint32_t b0 = __ballot_sync(0xFFFFFFFFu, tp), b1 = __ballot_sync(0xFFFFFFFFu, tn), b2 = __ballot_sync(0xFFFFFFFFu, fp);
    const uint32_t b3 = __ballot_sync(0xFFFFFFFFu, fn), b4 = __ballot_sync(0xFFFFFFFFu, se), b5 = __ballot_sync(0xFFFFFFFFu, dc);
    if(threadIdx.x == 0) {
        if(b0) atomicAdd(&s_cnt[BC_TP], __popc(b0)); if(b1) atomicAdd(&s_cnt[BC_TN], __popc(b1)); if(b2) atomicAdd(&s_cnt[BC_FP], __popc(b2));
        if(b3) atomicAdd(&s_cnt[BC_FN], __popc(b3)); if(b4) atomicAdd(&s_cnt[BC_SE], __popc(b4)); if(b5) atomicAdd(&s_cnt[BC_DC], __popc(b5));
    }
    __syncthreads();
    if(tid < BC_COUNT && s_cnt[tid]) atomicAdd(&A.counters[tid], (unsigned long long)s_cnt[tid]);
}

} // namespace lvb
