#include "../litiv_b200/csrc/common.cuh"
#include <cstdio>
#include <cstring>
#include <vector>
using namespace lvb;
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__global__ void k(const __grid_constant__ CUtensorMap tmap, int c0, int c1, int bytes, unsigned char* out) {
    __shared__ __align__(128) unsigned char tile[256 * 16];
    __shared__ __align__(8) uint64_t bar;
    if(threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    __syncthreads();
    if(threadIdx.x == 0) { mbar_expect_tx(&bar, bytes); tma_load_2d(tile, &tmap, c0, c1, &bar); }
    mbar_wait(&bar, 0);
    for(int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = tile[i];
}
int main() {
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    EncodeTiledFn fn = (EncodeTiledFn)p;
    const int W = 480, H = 120, pitch = 512;
    unsigned char* d; cudaMalloc(&d, pitch * H); 
    std::vector<unsigned char> h(pitch * H); for(size_t i = 0; i < h.size(); ++i) h[i] = (unsigned char)(i * 7 + i / pitch);
    cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    unsigned char* o; cudaMalloc(&o, 4096);
    struct T { int bx, by, c0, c1; CUtensorMapL2promotion l2; } tests[] = {
        {128, 12, 80, 6, CU_TENSOR_MAP_L2_PROMOTION_NONE}, {128, 12, -16, -2, CU_TENSOR_MAP_L2_PROMOTION_NONE}, {128, 12, 464, 110, CU_TENSOR_MAP_L2_PROMOTION_NONE},
        {64, 12, -16, -2, CU_TENSOR_MAP_L2_PROMOTION_NONE}, {64, 12, 464, 115, CU_TENSOR_MAP_L2_PROMOTION_L2_128B}, {128, 12, 96, 6, CU_TENSOR_MAP_L2_PROMOTION_L2_256B}};
    for(auto& t : tests) {
        CUtensorMap m; memset(&m, 0, sizeof(m));
        cuuint64_t gd[2] = {(cuuint64_t)W, (cuuint64_t)H}; cuuint64_t gs[1] = {(cuuint64_t)pitch};
        cuuint32_t bx[2] = {(cuuint32_t)t.bx, (cuuint32_t)t.by}; cuuint32_t es[2] = {1, 1};
        CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, t.l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        cudaMemset(o, 0xEE, 4096);
        k<<<1, 128>>>(m, t.c0, t.c1, t.bx * t.by, o);
        cudaError_t e = cudaDeviceSynchronize();
        std::vector<unsigned char> ho(4096); cudaMemcpy(ho.data(), o, 4096, cudaMemcpyDeviceToHost);
        int bad = 0;
        for(int r2 = 0; r2 < t.by; ++r2) for(int b = 0; b < t.bx; ++b) {
            int gy = t.c1 + r2, gx = t.c0 + b; unsigned char want = (gy >= 0 && gy < H && gx >= 0 && gx < W) ? h[gy * pitch + gx] : 0;
            bad += ho[r2 * t.bx + b] != want;
        }
        printf("box %dx%d at (%d,%d): encode=%d run=%s mismatches=%d\n", t.bx, t.by, t.c0, t.c1, (int)r, cudaGetErrorString(e), bad);
        if(e != cudaSuccess) { printf("sticky error, stop\n"); return 1; }
    }
    return 0;
}
