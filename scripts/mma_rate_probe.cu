// mma_rate_probe.cu — how fast does the B200 tensor core retire small tcgen05.mma (kind::tf32, M = 128, K = 8, SS operands)?
// One CTA per SM (148), one convergent warp issues NMMA instructions back to back, then commits and waits; clock64 around it.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mma_rate_probe scripts/mma_rate_probe.cu && /tmp/mma_rate_probe
// Variables: N (16..256), operand layout of A (0 = no swizzle interleaved, 4 = SWIZZLE_64B, 2 = SWIZZLE_128B), number of
// distinct accumulators cycled through (1 = every MMA depends on the previous one's accumulator), A start-address stride
// between MMAs (0 = same tile every time).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint64_t layout) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= layout << 61;
    return d;
}
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

template <int NMMA>
__global__ void __launch_bounds__(128, 1) probe(int N, int layout, int nacc, int astride16, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 1.0f;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = __shfl_sync(0xffffffffu, slot, 0);
    if (warp == 0) {
        // A: 128 rows; layout 0: [2 K-halves][128 rows][16 B] (LBO 2048, SBO 128); swizzled: rows 64 / 128 B apart
        const uint64_t ad0 = layout == 0 ? desc(smem_u32(smem), 2048, 128, 0)
                                         : desc(smem_u32(smem), 16, layout == 4 ? 8 * 64 : 8 * 128, (uint64_t)layout);
        const uint64_t bd = desc(smem_u32(smem) + 64 * 1024, (uint32_t)N * 16, 128, 0);     // B: [2][N rows][16 B]
        const uint32_t id = idesc_tf32(128, N);
        const long long t0 = clock64();
#pragma unroll 4
        for (int i = 0; i < NMMA; ++i) {
            const uint32_t dcol = tmem + (uint32_t)((i % nacc) * N);
            const uint64_t ad = ad0 + (uint64_t)((i & 15) * astride16);
            asm volatile(
                "{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                ::"r"(dcol), "l"(ad), "l"(bd), "r"(id), "r"(i >= nacc ? 1u : 0u) : "memory");
        }
        const long long t1 = clock64();
        asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
                     "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
        const long long t2 = clock64();
        if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

int main() {
    long long* out;
    cudaMallocManaged(&out, 16);
    constexpr int NMMA = 2048;
    cudaFuncSetAttribute(probe<NMMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    printf("N layout nacc astride issue_clk/mma total_clk/mma\n");
    for (int N : {16, 32, 48, 96, 128, 256})
        for (int layout : {0, 4, 2})
            for (int nacc : {1, 2})
                for (int astride : {0, 4}) {
                    if (nacc * N > 512) continue;
                    probe<NMMA><<<148, 128, 100 * 1024>>>(N, layout, nacc, astride, out);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("N=%d layout=%d: %s\n", N, layout, cudaGetErrorString(e)); return 1; }
                    printf("%3d %d %d %d  %7.1f %7.1f\n", N, layout, nacc, astride, out[0] / (double)NMMA, out[1] / (double)NMMA);
                }
    return 0;
}
