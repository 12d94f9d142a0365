// Standalone TMA probe (debug tool, not part of the library): loads one box with
// cp.async.bulk.tensor in several variants and checks the result.  usage: tma_probe <rank> <issue_mode> <bw>
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int RANK>
__global__ void probe(const __grid_constant__ CUtensorMap tmap, float* out, int nfloats, int mode, int c0in, int c1in, int flags, const float* fsrc) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    float* tile = reinterpret_cast<float*>(smem);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    __syncthreads();
    int c0 = c0in, c1 = c1in;
    if (flags & 4) {   // coordinates derived from per-thread float data (non-uniform provenance, like the real kernel)
        float a = fsrc[threadIdx.x], b = fsrc[threadIdx.x + 128];
        for (int o = 16; o > 0; o >>= 1) { a = fminf(a, __shfl_xor_sync(0xffffffffu, a, o)); b = fminf(b, __shfl_xor_sync(0xffffffffu, b, o)); }
        c0 = (int)floorf(a); c1 = (int)floorf(b);
    }
    bool issue = false;
    if (mode == 0) issue = threadIdx.x == 0;
    if (mode == 1 && threadIdx.x < 32) {
        uint32_t pred;
        asm volatile("{\n .reg .pred p;\n elect.sync _|p, 0xffffffff;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
        issue = pred != 0;
    }
    if (issue) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar)), "r"(nfloats * 4));
        if (RANK == 2)
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                         ::"r"(s32(tile)), "l"(&tmap), "r"(s32(&bar)), "r"(c0), "r"(c1) : "memory");
        if (RANK == 3)
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                         ::"r"(s32(tile)), "l"(&tmap), "r"(s32(&bar)), "r"(c0), "r"(c1), "r"(0) : "memory");
        if (RANK == 5 && !(flags & 1))
            asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                         ::"r"(s32(tile)), "l"(&tmap), "r"(s32(&bar)), "r"(c0), "r"(c1), "r"(0), "r"(0), "r"(1) : "memory");
        if (RANK == 5 && (flags & 1))
            asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                         ::"r"(s32(tile)), "l"(&tmap), "r"(s32(&bar)), "r"(c0), "r"(c1), "r"(0), "r"(0), "r"(1) : "memory");
    }
    uint32_t done = 0, spins = 0;
    while (!done) {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(s32(&bar)) : "memory");
        if (!done && ++spins > (1u << 22)) { if (threadIdx.x == 0) printf("TIMEOUT waiting for TMA\n"); return; }
    }
    for (int i = threadIdx.x; i < nfloats; i += blockDim.x) out[i] = tile[i];
}

int main(int argc, char** argv) {
    const int rank = argc > 1 ? atoi(argv[1]) : 2, mode = argc > 2 ? atoi(argv[2]) : 0, bw = argc > 3 ? atoi(argv[3]) : 48, flags = argc > 4 ? atoi(argv[4]) : 0;
    const int W = argc > 5 ? atoi(argv[5]) : 192, H = argc > 6 ? atoi(argv[6]) : 128, CPG = argc > 7 ? atoi(argv[7]) : 2, G = 8, BV = 3;
    const int bh = argc > 8 ? atoi(argv[8]) : 16, bcp = argc > 9 ? atoi(argv[9]) : CPG, bg = argc > 10 ? atoi(argv[10]) : G;
    const size_t n = (size_t)W * H * CPG * G * BV;
    std::vector<float> h(n);
    for (size_t i = 0; i < n; ++i) h[i] = (float)(i % 100003);
    float *d, *out;
    cudaMalloc(&d, n * 4);
    cudaMemcpy(d, h.data(), n * 4, cudaMemcpyHostToDevice);
    void* fnp = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
    EncodeTiledFn fn = (EncodeTiledFn)fnp;
    CUtensorMap map;
    const cuuint64_t hw = (cuuint64_t)H * W;
    cuuint64_t dims[5] = {W, H, CPG, G, BV};
    cuuint64_t strides[4] = {(cuuint64_t)W * 4, hw * 4, hw * 4 * CPG, hw * 4 * CPG * G};
    cuuint32_t box[5] = {(cuuint32_t)bw, (cuuint32_t)bh, 1, 1, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    int cc = 1;
    if (rank == 5) { box[2] = bcp; box[3] = bg; cc = bcp * bg; }
    if (rank == 2) dims[1] = (cuuint64_t)H * CPG * G * BV;
    if (rank == 3) { dims[2] = (cuuint64_t)CPG * G * BV; }
    CUresult rc = fn(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("rank %d mode %d bw %d bh %d W %d H %d CPG %d box_cp %d box_g %d: encode rc=%d\n", rank, mode, bw, bh, W, H, CPG, bcp, bg, (int)rc);
    const int nfl = bw * bh * cc;
    cudaMalloc(&out, nfl * 4);
    cudaMemset(out, 0, nfl * 4);
    const int c0 = argc > 11 ? atoi(argv[11]) : ((flags & 2) ? -5 : 8), c1 = argc > 12 ? atoi(argv[12]) : ((flags & 2) ? -3 : 4);
    printf("  c0 %d c1 %d", c0, c1);
    std::vector<float> fs(256);
    for (int i = 0; i < 128; ++i) { fs[i] = c0 + 0.25f + (i % 7); fs[128 + i] = c1 + 0.5f + (i % 5); }
    float* fsrc; cudaMalloc(&fsrc, 1024); cudaMemcpy(fsrc, fs.data(), 1024, cudaMemcpyHostToDevice);
    printf("  flags %d\n", flags);
    size_t smem = (size_t)nfl * 4;
    cudaError_t e;
    if (rank == 2) { cudaFuncSetAttribute(probe<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); probe<2><<<1, 128, smem>>>(map, out, nfl, mode, c0, c1, flags, fsrc); }
    if (rank == 3) { cudaFuncSetAttribute(probe<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); probe<3><<<1, 128, smem>>>(map, out, nfl, mode, c0, c1, flags, fsrc); }
    if (rank == 5) { cudaFuncSetAttribute(probe<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); probe<5><<<1, 128, smem>>>(map, out, nfl, mode, c0, c1, flags, fsrc); }
    e = cudaDeviceSynchronize();
    printf("  sync: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<float> o(nfl);
    cudaMemcpy(o.data(), out, nfl * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    const int view = rank == 5 ? 1 : 0;
    for (int c = 0; c < cc; ++c)
        for (int y = 0; y < bh; ++y)
            for (int x = 0; x < bw; ++x) {
                const int yy = c1 + y, xx = c0 + x;
                const int cg = c / bcp, ccp = c % bcp;
                const float want = (yy < 0 || xx < 0 || yy >= H || xx >= W) ? 0.0f : h[((size_t)view * CPG * G + cg * CPG + ccp) * hw + (size_t)yy * W + xx];
                if (o[(c * bh + y) * bw + x] != want) ++bad;
            }
    printf("  mismatches: %d of %d\n", bad, nfl);
    return 0;
}
