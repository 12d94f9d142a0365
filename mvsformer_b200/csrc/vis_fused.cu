// vis_fused.cu — the visibility net (models/mvsformer_model.py:37,91: ConvBnReLU 1->16, 16->16, 16->8, Conv2d 8->1, Sigmoid) as
// ONE persistent kernel: entropy map in, visibility weight out, nothing else touches HBM.
//
// Why (profiles/r02c_launches.csv): as four kernels the net moved its [maps,H,W,16] activations through HBM three times —
// 2.7 GB per reference view at stage 4, more than the whole cost-volume build's byte budget — for 1.13 ms of the 5.3 ms step.
//
// A CTA lives for the whole launch and takes tiles of 30 x 14 output pixels of one map.  Per tile:
//   producer warps   layer 1 (1->16, 3x3) on CUDA cores for the 34 x 18 pixels layer 2 needs, written straight into shared memory
//                    as layer 2's tensor-core operand: channels-last 64-byte pixels in the 64-byte-swizzled K-major layout
//                    (the same layout TMA produces in conv3d_tma.cu; here the threads apply the address swizzle themselves);
//   MMA warp         layer 2 (16->16) as 4 M-tiles x 9 taps x 2 tcgen05.mma (kind::tf32, M = 128 = 16 rows x 8 pixels, N = 16),
//                    a tap being a pixel shift of the descriptor start address — accumulators in TMEM;
//   epilogue warps   drain layer 2 (+ folded-BN shift, ReLU, TF32 round, zero outside the image = layer 3's padding) back into
//                    shared memory as layer 3's operand; after layer 3's MMAs (16->8, N padded to 16) drain again and finish
//                    in registers: ReLU, the 1x1 conv 8->1, sigmoid, one float per pixel to HBM.
// Layer 2 of tile i+1 is issued before layer 3 of tile i (double-buffered layer-1 operand and layer-2 accumulators), so the
// tensor core, the CUDA-core producers and the epilogue warps overlap.  Arithmetic matches the unfused kernels operation by
// operation (same FMA order in layers 1 and 4, same TF32 operands and rounding points), so the result is bit-identical.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "tcgen05.cuh"

namespace mvs {
namespace tc {
namespace visf {

constexpr int OUT_Y = 30, OUT_X = 14;            // output pixels per tile
constexpr int RY = 32, RX = 16;                  // layer-2 region = 2 x 2 M-tiles of 16 rows x 8 pixels
constexpr int AY = 34, AX = 18;                  // operand arrays: region + 1-pixel halo
constexpr int PIXB = 64;                         // bytes per pixel (16 channels)
constexpr int ABYTES = (AY * AX * PIXB + 1023) / 1024 * 1024;
constexpr int WBYTES = 9 * 4 * 16 * 16;          // one packed 3x3 16->16 weight tile
constexpr int NPROD = 4, NEPI = 8;               // producer / epilogue warps
constexpr int THREADS = 32 * (1 + NPROD + NEPI);
constexpr size_t SMEM = 3 * ABYTES + 2 * WBYTES + 256 + 1024;

struct Params {
    float w1[16][9]; float b1[16];               // layer 1, BN folded
    float shift2[16]; float shift3[8];           // folded BN shifts of layers 2 and 3
    float w4[8]; float b4;                       // 1x1 conv
};

struct Dims { int M, H, W, tiles_x, tiles_y, nitems; };

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mma_tf32_ss_elect(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p, q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void mma_commit_elect(uint64_t* bar) {
    asm volatile(
        "{\n\t"
        ".reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t"
        "}" ::"r"(smem_u32(bar)) : "memory");
}
// K-major operand with the 64-byte swizzle: rows (pixels) 64 B apart, 8-row groups `sbo` bytes apart
__device__ __forceinline__ uint64_t make_desc_sw64(uint32_t saddr, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)4 << 61;
    return d;
}
// byte offset of 16-byte chunk c of pixel `pix` inside a 1024-byte-aligned operand array: Swizzle<2,4,3> on the address
__device__ __forceinline__ uint32_t sw64_offset(int pix, int c) {
    const uint32_t off = (uint32_t)pix * PIXB;
    return off + (uint32_t)((c ^ ((off >> 7) & 3)) << 4);
}

__global__ void __launch_bounds__(THREADS, 1)
vis_fused_kernel(const float* __restrict__ entropy, const float* __restrict__ w2, const float* __restrict__ w3,
                 float* __restrict__ weight, Dims d, const __grid_constant__ Params P) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sA1 = smem;                                   // [2][AY][AX][64 B]  layer-1 output = layer-2 operand
    uint8_t* sA2 = smem + 2 * ABYTES;                      // [AY][AX][64 B]     layer-2 output = layer-3 operand (border stays zero)
    uint8_t* sW = sA2 + ABYTES;                            // W2, W3 packed [tap][quad][16 n][4]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sW + 2 * WBYTES);
    uint64_t *full1 = bars, *free1 = bars + 2, *acc2_full = bars + 4, *acc2_free = bars + 6;
    uint64_t *full2 = bars + 8, *free2 = bars + 9, *acc3_full = bars + 10, *acc3_free = bars + 11;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

    // warp index broadcast from lane 0: the compiler then knows it is warp-uniform, keeps the role branches and everything
    // derived inside them (descriptors, TMEM addresses) on the uniform datapath — the MMA issue loop is the pipeline's pace-maker
    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&full1[i], NPROD); mbar_init(&free1[i], 1);
            mbar_init(&acc2_full[i], 1); mbar_init(&acc2_free[i], NEPI * 32);
        }
        mbar_init(full2, NEPI * 32); mbar_init(free2, 1); mbar_init(acc3_full, 1); mbar_init(acc3_free, NEPI * 32);
        mbar_fence_init();
    }
    // weights -> shared memory (plain copies; made visible to the tensor core by the proxy fence below), A2 zeroed once
    for (int i = tid; i < 2 * WBYTES / 16; i += THREADS)
        reinterpret_cast<float4*>(sW)[i] = __ldg(reinterpret_cast<const float4*>(i < WBYTES / 16 ? w2 : w3) + (i < WBYTES / 16 ? i : i - WBYTES / 16));
    for (int i = tid; i < ABYTES / 16; i += THREADS) reinterpret_cast<float4*>(sA2)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (warp == 0) tmem_alloc(tmem_slot, 256);
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);

    if (warp == 0) {
        // ================================================= MMA issuer (whole warp, convergent) =========================
        constexpr uint32_t idesc = make_idesc_tf32(128, 16);
        const uint64_t a1d = make_desc_sw64(smem_u32(sA1), AX * PIXB), a2d = make_desc_sw64(smem_u32(sA2), AX * PIXB);
        const uint64_t b2d = make_smem_desc(smem_u32(sW), 256, 128), b3d = make_smem_desc(smem_u32(sW) + WBYTES, 256, 128);
        // one layer over the four M-tiles of a region: 9 taps x 2 K chunks each, a tap = a pixel shift of the start address
        auto layer = [&](uint64_t ad, uint64_t bd, uint32_t acc0) {
#pragma unroll 1
            for (int t = 0; t < 4; ++t) {
                const uint64_t at = ad + (uint64_t)((((t >> 1) * 16 * AX + (t & 1) * 8) * PIXB) >> 4);
                const uint32_t dcol = acc0 + (uint32_t)t * 16;
#pragma unroll
                for (int tap = 0; tap < 9; ++tap)
#pragma unroll
                    for (int kk = 0; kk < 2; ++kk)
                        mma_tf32_ss_elect(dcol, at + (uint64_t)((((tap / 3) * AX + (tap % 3)) * PIXB + kk * 32) >> 4),
                                          bd + (uint64_t)((tap * 4 + 2 * kk) * 16), idesc, (tap | kk) ? 1u : 0u);
            }
        };
        auto layer2 = [&](int j) {
            const int b = j & 1;
            mbar_wait(&full1[b], (j >> 1) & 1);
            if (j >= 2) mbar_wait(&acc2_free[b], ((j >> 1) - 1) & 1);
            tc_fence_after_sync();
            layer(a1d + (uint64_t)(b * (ABYTES >> 4)), b2d, tmem + (uint32_t)b * 64);
            mma_commit_elect(&free1[b]);
            mma_commit_elect(&acc2_full[b]);
        };
        int it = 0;
        if ((int)blockIdx.x < d.nitems) layer2(0);
        for (int item = blockIdx.x; item < d.nitems; item += gridDim.x, ++it) {
            if (item + (int)gridDim.x < d.nitems) layer2(it + 1);
            mbar_wait(full2, it & 1);
            if (it >= 1) mbar_wait(acc3_free, (it - 1) & 1);
            tc_fence_after_sync();
            layer(a2d, b3d, tmem + 128);
            mma_commit_elect(free2);
            mma_commit_elect(acc3_full);
        }
    } else if (warp <= NPROD) {
        // ================================================= layer-1 producers (CUDA cores) ==============================
        const int pt = (warp - 1) * 32 + lane;
        int it = 0;
        for (int item = blockIdx.x; item < d.nitems; item += gridDim.x, ++it) {
            const int b = it & 1;
            const int tx = item % d.tiles_x, ty = (item / d.tiles_x) % d.tiles_y, m = item / (d.tiles_x * d.tiles_y);
            const float* ent = entropy + (int64_t)m * d.H * d.W;
            if (it >= 2) mbar_wait(&free1[b], ((it >> 1) - 1) & 1);
            uint8_t* a1 = sA1 + b * ABYTES;
            for (int pix = pt; pix < AY * AX; pix += NPROD * 32) {
                const int y = ty * OUT_Y - 2 + pix / AX, x = tx * OUT_X - 2 + pix % AX;
                float r[16];
                if (y >= 0 && y < d.H && x >= 0 && x < d.W) {
                    float in[9];
#pragma unroll
                    for (int t = 0; t < 9; ++t) {
                        const int yy = y - 1 + t / 3, xx = x - 1 + t % 3;
                        in[t] = (yy >= 0 && yy < d.H && xx >= 0 && xx < d.W) ? __ldg(ent + (int64_t)yy * d.W + xx) : 0.0f;
                    }
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        float a = P.b1[e];
#pragma unroll
                        for (int t = 0; t < 9; ++t) a = fmaf(in[t], P.w1[e][t], a);
                        r[e] = round_to_tf32(fmaxf(a, 0.0f));
                    }
                } else {
#pragma unroll
                    for (int e = 0; e < 16; ++e) r[e] = 0.0f;     // outside the image: layer 2's zero padding
                }
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    *reinterpret_cast<float4*>(a1 + sw64_offset(pix, c)) = make_float4(r[4 * c], r[4 * c + 1], r[4 * c + 2], r[4 * c + 3]);
            }
            fence_proxy_async_smem();                      // generic-proxy stores -> visible to the tensor core
            __syncwarp();
            if (lane == 0) mbar_arrive(&full1[b]);
        }
    } else {
        // ================================================= epilogue warps ==============================================
        const int q = warp & 3;                            // TMEM lane quarter this warp may read
        const int half = (warp - 1 - NPROD) >> 2;          // M-tiles t = half, half + 2
        const int mrow = q * 32 + lane;                    // accumulator row = pixel of the M-tile
        const int tyl = mrow >> 3, txl = mrow & 7;
        const uint32_t tlane = tmem + ((uint32_t)(q * 32) << 16);
        int it = 0;
        for (int item = blockIdx.x; item < d.nitems; item += gridDim.x, ++it) {
            const int b = it & 1;
            const int tx = item % d.tiles_x, ty = (item / d.tiles_x) % d.tiles_y, m = item / (d.tiles_x * d.tiles_y);
            // ---- layer 2 -> operand of layer 3
            mbar_wait(&acc2_full[b], (it >> 1) & 1);
            if (it >= 1) mbar_wait(free2, (it - 1) & 1);   // layer 3 of the previous tile has read A2
            tc_fence_after_sync();
#pragma unroll
            for (int tt = 0; tt < 2; ++tt) {
                const int t = half + 2 * tt;
                float acc[16];
                tmem_ld16(tlane + (uint32_t)b * 64 + (uint32_t)t * 16, acc);
                const int ry = (t >> 1) * 16 + tyl, rx = (t & 1) * 8 + txl;           // position in the layer-2 region
                const int y = ty * OUT_Y - 1 + ry, x = tx * OUT_X - 1 + rx;
                const bool inside = y >= 0 && y < d.H && x >= 0 && x < d.W;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    float4 v;
                    v.x = inside ? round_to_tf32(fmaxf(acc[4 * c] + P.shift2[4 * c], 0.f)) : 0.f;
                    v.y = inside ? round_to_tf32(fmaxf(acc[4 * c + 1] + P.shift2[4 * c + 1], 0.f)) : 0.f;
                    v.z = inside ? round_to_tf32(fmaxf(acc[4 * c + 2] + P.shift2[4 * c + 2], 0.f)) : 0.f;
                    v.w = inside ? round_to_tf32(fmaxf(acc[4 * c + 3] + P.shift2[4 * c + 3], 0.f)) : 0.f;
                    *reinterpret_cast<float4*>(sA2 + sw64_offset(ry * AX + rx, c)) = v;
                }
            }
            fence_proxy_async_smem();
            tc_fence_before_sync();
            mbar_arrive(&acc2_free[b]);
            mbar_arrive(full2);
            // ---- layer 3 -> ReLU -> 1x1 conv -> sigmoid
            mbar_wait(acc3_full, it & 1);
            tc_fence_after_sync();
#pragma unroll
            for (int tt = 0; tt < 2; ++tt) {
                const int t = half + 2 * tt;
                float acc[16];
                tmem_ld16(tlane + 128 + (uint32_t)t * 16, acc);
                const int oy_l = (t >> 1) * 16 + tyl, ox_l = (t & 1) * 8 + txl;       // output pixel of the tile
                const int y = ty * OUT_Y + oy_l, x = tx * OUT_X + ox_l;
                if (oy_l < OUT_Y && ox_l < OUT_X && y < d.H && x < d.W) {
                    float z = P.b4;
#pragma unroll
                    for (int i = 0; i < 8; ++i) z = fmaf(round_to_tf32(fmaxf(acc[i] + P.shift3[i], 0.f)), P.w4[i], z);
                    weight[((int64_t)m * d.H + y) * d.W + x] = 1.0f / (1.0f + expf(-z));
                }
            }
            tc_fence_before_sync();
            mbar_arrive(acc3_free);
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 256);
}

}  // namespace visf
}  // namespace tc
}  // namespace mvs

// entropy [M,H,W] -> visibility weight [M,H,W].  params_host: w1[16][9] b1[16] shift2[16] shift3[8] w4[8] b4 (BN folded);
// w2 / w3: device, packed like mvs_conv3d_tma weights for kd = 1, Cin = 16, n_tile = 16 ([kh][kw][4 quads][16 n][4], TF32).
extern "C" int mvs_vis_fused(const float* entropy, const float* params_host, const float* w2, const float* w3, float* weight,
                             int M, int H, int W, void* stream) {
    using namespace mvs::tc::visf;
    MVS_REQUIRE(entropy && params_host && w2 && w3 && weight, "mvs_vis_fused: null pointer");
    MVS_REQUIRE(M >= 1 && H >= 1 && W >= 1, "mvs_vis_fused: empty shape M=%d H=%d W=%d", M, H, W);
    MVS_REQUIRE((((uintptr_t)w2 | (uintptr_t)w3) & 15) == 0, "mvs_vis_fused: packed weights must be 16-byte aligned");
    Params P;
    memcpy(&P, params_host, sizeof(P));
    Dims d{M, H, W, mvs::cdiv(W, OUT_X), mvs::cdiv(H, OUT_Y), 0};
    d.nitems = M * d.tiles_x * d.tiles_y;
    static int nsm = 0;
    if (!nsm) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
        if (nsm <= 0) nsm = 148;
    }
    MVS_CUDA_OK(cudaFuncSetAttribute(vis_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
    const int grid = d.nitems < nsm ? d.nitems : nsm;
    vis_fused_kernel<<<grid, THREADS, SMEM, (cudaStream_t)stream>>>(entropy, w2, w3, weight, d, P);
    MVS_LAUNCH_OK("vis_fused_kernel");
    return MVS_OK;
}
