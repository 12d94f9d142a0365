// vis_fused.cu — the visibility net (models/mvsformer_model.py:37,91: ConvBnReLU 1->16, 16->16, 16->8, Conv2d 8->1, Sigmoid) as
// ONE persistent kernel: entropy map in, visibility weight out, nothing else touches HBM.
//
// Why (profiles/r02c_launches.csv): as four kernels the net moved its [maps,H,W,16] activations through HBM three times —
// 2.7 GB per reference view at stage 4, more than the whole cost-volume build's byte budget — for 1.13 ms of the 5.3 ms step.
//
// A CTA lives for the whole launch and takes tiles of 28 x 14 output pixels of one map.  Per tile:
//   producer warps   layer 1 (1->16, 3x3) on CUDA cores for the 32 x 18 pixels layer 2 reads, written straight into shared memory
//                    as layer 2's tensor-core operand: channels-last 64-byte pixels in the 64-byte-swizzled K-major layout
//                    (the same layout TMA produces in conv3d_tma.cu; here the threads apply the address swizzle themselves);
//   MMA warp         layers 2 (16->16) and 3 (16->8) as tcgen05.mma kind::tf32 with the three kh taps FUSED INTO N: an M-tile is
//                    16 INPUT rows x 8 pixels, the B operand of a (kw, K chunk) is [W(kh=0) | W(kh=1) | W(kh=2)], so one MMA
//                    yields the partial sums T[kh](input row) of all three kernel rows: 6 MMAs per M-tile and layer instead
//                    of 18.  (Measured, scripts/mma_rate_probe.cu: a tf32 M = 128 MMA costs ~75 clk for ANY N <= 128, so
//                    N = 48 is as cheap as N = 16 — the first version of this kernel, 18 MMAs of N = 16, sat at exactly that
//                    issue floor: 0.75 ms.)
//   epilogue warps   drain the accumulators, exchange T[1], T[2] through shared memory and form
//                    out(y) = T[0](y) + T[1](y+1) + T[2](y+2); layer 2: + folded-BN shift, ReLU, TF32 round, zero outside the
//                    image (= layer 3's padding), written back to shared memory as layer 3's operand; layer 3: ReLU, the 1x1
//                    conv 8->1, sigmoid, one float per pixel to HBM.
// Layer 2 of tile i+1 is issued before layer 3 of tile i (double-buffered layer-1 operand and layer-2 accumulators), so the
// tensor core, the CUDA-core producers and the epilogue warps overlap.  Operands and rounding points are those of the
// unfused kernels; only the order of the three per-kh partial sums differs (fp32, ~1e-7 relative).
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "tcgen05.cuh"

namespace mvs {
namespace tc {
namespace visf {

constexpr int OUT_Y = 28, OUT_X = 14;            // output pixels per tile
constexpr int AY = 32, AX = 18;                  // operand arrays: 2 x 2 M-tiles of 16 input rows x 8 pixels (+2 columns of kw halo)
constexpr int PIXB = 64;                         // bytes per pixel (16 channels)
constexpr int ABYTES = AY * AX * PIXB;           // 36 KB, a multiple of 1024
constexpr int N2 = 48, N3 = 32;                  // MMA N: [3 kh][16] and [3 kh][8] padded to 32
constexpr int W2BYTES = 3 * 4 * N2 * 16, W3BYTES = 3 * 4 * N3 * 16;      // packed [kw][quad][kh][n][4]
constexpr int XBYTES = 2 * AY * 16 * PIXB;       // exchange buffers for T[1], T[2]: [2][32 rows][16 pixels][64 B]
constexpr int NPROD = 7, NEPI = 8;               // producer / epilogue warps
constexpr int THREADS = 32 * (1 + NPROD + NEPI);
constexpr size_t SMEM = 3 * ABYTES + W2BYTES + W3BYTES + XBYTES + 256 + 1024;
static_assert(ABYTES % 1024 == 0, "operand arrays keep the swizzle phase");

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
__device__ __forceinline__ void epi_sync() { asm volatile("bar.sync 1, %0;" ::"n"(NEPI * 32) : "memory"); }   // epilogue warps only
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
// byte offset of 16-byte chunk c of pixel `pix` inside a 1024-byte-aligned array of 64-byte pixels: Swizzle<2,4,3> on the address
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
    uint8_t* sA2 = smem + 2 * ABYTES;                      // [AY][AX][64 B]     layer-2 output = layer-3 operand (rest stays zero)
    uint8_t* sX = sA2 + ABYTES;                            // [2][AY][16][64 B]  T[1], T[2] exchange
    uint8_t* sW2 = sX + XBYTES;                            // [kw][quad][3 kh x 16 n][4]
    uint8_t* sW3 = sW2 + W2BYTES;                          // [kw][quad][3 kh x 8 n + 8 pad][4]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sW3 + W3BYTES);
    uint64_t *full1 = bars, *free1 = bars + 2, *acc2_full = bars + 4, *acc2_free = bars + 5;
    uint64_t *full2 = bars + 6, *free2 = bars + 7, *acc3_full = bars + 8, *acc3_free = bars + 10;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

    // warp index broadcast from lane 0: the compiler then knows it is warp-uniform, keeps the role branches and everything
    // derived inside them (descriptors, TMEM addresses) on the uniform datapath — the MMA issue loop is the pipeline's pace-maker
    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&full1[i], NPROD); mbar_init(&free1[i], 1);
            mbar_init(&acc3_full[i], 1); mbar_init(&acc3_free[i], NEPI * 32);
        }
        mbar_init(full2, NEPI * 32); mbar_init(free2, 1); mbar_init(acc2_full, 1); mbar_init(acc2_free, NEPI * 32);
        mbar_fence_init();
    }
    // weights -> shared memory (plain copies; made visible to the tensor core by the proxy fence below), A2 zeroed once
    for (int i = tid; i < W2BYTES / 16; i += THREADS) reinterpret_cast<float4*>(sW2)[i] = __ldg(reinterpret_cast<const float4*>(w2) + i);
    for (int i = tid; i < W3BYTES / 16; i += THREADS) reinterpret_cast<float4*>(sW3)[i] = __ldg(reinterpret_cast<const float4*>(w3) + i);
    for (int i = tid; i < ABYTES / 16; i += THREADS) reinterpret_cast<float4*>(sA2)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (warp == 0) tmem_alloc(tmem_slot, 512);
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    // TMEM columns: layer-2 accumulators [4 M-tiles][48] (drained right away, so one buffer), layer-3 accumulators
    // [2 buffers][4 M-tiles][32]: the epilogue finishes tile i-1 (layer 3) AFTER it has handed tile i's layer-2 output to
    // the tensor core, so layer 3 of tile i runs while tile i-1's result is still being drained
    constexpr uint32_t ACC3 = 4 * N2;

    if (warp == 0) {
        // ================================================= MMA issuer (whole warp, convergent) =========================
        const uint64_t a1d = make_desc_sw64(smem_u32(sA1), AX * PIXB), a2d = make_desc_sw64(smem_u32(sA2), AX * PIXB);
        const uint64_t b2d = make_smem_desc(smem_u32(sW2), N2 * 16, 128), b3d = make_smem_desc(smem_u32(sW3), N3 * 16, 128);
        // one layer over the four M-tiles: 3 kw x 2 K chunks, N = the three kh taps; a kw tap = a pixel shift of the start address
        auto layer = [&](uint64_t ad, uint64_t bd, uint32_t acc0, uint32_t n, uint32_t idesc) {
#pragma unroll 1
            for (int t = 0; t < 4; ++t) {
                const uint64_t at = ad + (uint64_t)((((t >> 1) * 16 * AX + (t & 1) * 8) * PIXB) >> 4);
                const uint32_t dcol = acc0 + (uint32_t)t * n;
#pragma unroll
                for (int kw = 0; kw < 3; ++kw)
#pragma unroll
                    for (int kk = 0; kk < 2; ++kk)
                        mma_tf32_ss_elect(dcol, at + (uint64_t)((kw * PIXB + kk * 32) >> 4), bd + (uint64_t)((kw * 4 + 2 * kk) * n), idesc,
                                          (kw | kk) ? 1u : 0u);
            }
        };
        auto layer2 = [&](int j) {
            const int b = j & 1;
            mbar_wait(&full1[b], (j >> 1) & 1);
            if (j >= 1) mbar_wait(acc2_free, (j - 1) & 1);
            tc_fence_after_sync();
            layer(a1d + (uint64_t)(b * (ABYTES >> 4)), b2d, tmem, N2, make_idesc_tf32(128, N2));
            mma_commit_elect(&free1[b]);
            mma_commit_elect(acc2_full);
        };
        int it = 0;
        if ((int)blockIdx.x < d.nitems) layer2(0);
        for (int item = blockIdx.x; item < d.nitems; item += gridDim.x, ++it) {
            if (item + (int)gridDim.x < d.nitems) layer2(it + 1);
            const int b3 = it & 1;
            mbar_wait(full2, it & 1);
            if (it >= 2) mbar_wait(&acc3_free[b3], ((it >> 1) - 1) & 1);
            tc_fence_after_sync();
            layer(a2d, b3d, tmem + ACC3 + (uint32_t)b3 * 4 * N3, N3, make_idesc_tf32(128, N3));
            mma_commit_elect(free2);
            mma_commit_elect(&acc3_full[b3]);
        }
    } else if (warp <= NPROD) {
        // ================================================= layer-1 producers (CUDA cores) ==============================
        // A lane computes TWO vertically adjacent pixels of the 32 x 18 operand array at a time, one in each half of a
        // packed FP32 pair: a·w + acc for both pixels is ONE FFMA2 with the weight broadcast (half the FMA-pipe slots of
        // the scalar form; each half is the same IEEE fma, so the results are bit-identical), and the pair shares its
        // 4 x 3 entropy footprint (12 loads instead of 18).  Consecutive lanes take consecutive pixels of a row, so their
        // 64-byte operand rows fall on different banks (a horizontal pair per lane was measured 2-way conflicted: 128-byte
        // lane stride).  The 288 pairs of a tile are dealt to the 224 producer lanes starting at a warp that rotates with
        // the tile, so the 64 left-over pairs do not always land on the same warps; the footprint of a lane's first pair
        // of the NEXT tile is loaded as soon as this tile's has been consumed, and this tile's second pair before the
        // wait for the operand buffer: no global-load latency sits between the wait and the FMAs.
        constexpr int NPAIR = (AY / 2) * AX, PLANES = NPROD * 32, EB = 8;     // EB: channels per register block
        static_assert(AY % 2 == 0 && NPAIR > PLANES && NPAIR <= 2 * PLANES, "pair schedule");
        auto pair_of = [&](int it, int second) { return ((warp - 1 + it) % NPROD) * 32 + lane + second * PLANES; };
        struct Org { int y0, x0; const float* ent; };                                     // A1 origin (= tile origin - 2), map base
        auto origin = [&](int item) {
            const int tx = item % d.tiles_x, ty = (item / d.tiles_x) % d.tiles_y, m = item / (d.tiles_x * d.tiles_y);
            return Org{ty * OUT_Y - 2, tx * OUT_X - 2, entropy + (int64_t)m * d.H * d.W};
        };
        // entropy footprint rows y-1..y+2, columns x-1..x+1 of the pair at (y, x), (y+1, x); zero outside the image
        auto load12 = [&](const Org& o, int pr, float (&v)[12]) {
            const int y = o.y0 + 2 * (pr / AX), x = o.x0 + pr % AX;
            bool cok[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) cok[c] = (unsigned)(x - 1 + c) < (unsigned)d.W;
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int yy = y - 1 + r;
                const bool rok = (unsigned)yy < (unsigned)d.H;
                const float* row = o.ent + (int64_t)yy * d.W + (x - 1);
#pragma unroll
                for (int c = 0; c < 3; ++c) v[r * 3 + c] = (rok && cok[c]) ? __ldg(row + c) : 0.0f;
            }
        };
        auto compute_pair = [&](const Org& o, int pr, const float (&v)[12], uint8_t* a1) {
            const int py = 2 * (pr / AX), px = pr % AX;
            const int y = o.y0 + py, x = o.x0 + px;
            const bool colin = (unsigned)x < (unsigned)d.W;
            const bool in0 = colin && (unsigned)y < (unsigned)d.H, in1 = colin && (unsigned)(y + 1) < (unsigned)d.H;
            f32x2 in[9];                                                                  // tap t of (upper pixel, lower pixel)
#pragma unroll
            for (int t = 0; t < 9; ++t) in[t] = pack2(v[(t / 3) * 3 + t % 3], v[(t / 3 + 1) * 3 + t % 3]);
            const int pix = py * AX + px;
#pragma unroll
            for (int eh = 0; eh < 16; eh += EB) {                                         // EB channels at a time (registers)
                float r0[EB], r1[EB];
#pragma unroll
                for (int e = 0; e < EB; ++e) {
                    f32x2 a = pack2(P.b1[eh + e], P.b1[eh + e]);
#pragma unroll
                    for (int t = 0; t < 9; ++t) a = ffma2(in[t], pack2(P.w1[eh + e][t], P.w1[eh + e][t]), a);
                    const float2 o2 = unpack2(a);
                    r0[e] = in0 ? relu_round_tf32(o2.x) : 0.0f;                           // outside the image: layer 2's zero padding
                    r1[e] = in1 ? relu_round_tf32(o2.y) : 0.0f;
                }
#pragma unroll
                for (int c = 0; c < EB / 4; ++c) {
                    *reinterpret_cast<float4*>(a1 + sw64_offset(pix, eh / 4 + c)) = make_float4(r0[4 * c], r0[4 * c + 1], r0[4 * c + 2], r0[4 * c + 3]);
                    *reinterpret_cast<float4*>(a1 + sw64_offset(pix + AX, eh / 4 + c)) = make_float4(r1[4 * c], r1[4 * c + 1], r1[4 * c + 2], r1[4 * c + 3]);
                }
            }
        };
        int it = 0;
        float nxt[12];
        Org cur = origin(blockIdx.x);                      // (unused when the CTA has no tile)
        if ((int)blockIdx.x < d.nitems) load12(cur, pair_of(0, 0), nxt);
        for (int item = blockIdx.x; item < d.nitems; item += gridDim.x, ++it) {
            const int b = it & 1;
            const int p0 = pair_of(it, 0), p1 = pair_of(it, 1);
            const bool more = item + (int)gridDim.x < d.nitems;
            const Org nxo = origin(more ? item + (int)gridDim.x : item);
            float sec[12];
            if (p1 < NPAIR) load12(cur, p1, sec);
            if (it >= 2) mbar_wait(&free1[b], ((it >> 1) - 1) & 1);
            uint8_t* a1 = sA1 + b * ABYTES;
            compute_pair(cur, p0, nxt, a1);
            if (more) load12(nxo, pair_of(it + 1, 0), nxt);                               // lands under the next wait
            if (p1 < NPAIR) compute_pair(cur, p1, sec, a1);
            fence_proxy_async_smem();                      // generic-proxy stores -> visible to the tensor core
            __syncwarp();
            if (lane == 0) mbar_arrive(&full1[b]);
            cur = nxo;
        }
    } else {
        // ================================================= epilogue warps ==============================================
        const int q = warp & 3;                            // TMEM lane quarter this warp may read
        const int half = (warp - 1 - NPROD) >> 2;          // M-tiles t = half, half + 2
        const int mrow = q * 32 + lane;                    // accumulator row = (input row, pixel) of the M-tile
        const int tyl = mrow >> 3, txl = mrow & 7;
        const uint32_t tlane = tmem + ((uint32_t)(q * 32) << 16);
        uint8_t* sX1 = sX;                                 // T[1] by input row
        uint8_t* sX2 = sX + XBYTES / 2;                    // T[2] by input row
        // layer 3 of tile `jt` (item index `jitem`): out3(oy, ox) = T0(oy) + T1(oy + 1) + T2(oy + 2) -> ReLU -> 1x1 conv -> sigmoid
        auto finish = [&](int jt, int jitem) {
            const int b3 = jt & 1;
            const int tx = jitem % d.tiles_x, ty = (jitem / d.tiles_x) % d.tiles_y, m = jitem / (d.tiles_x * d.tiles_y);
            mbar_wait(&acc3_full[b3], (jt >> 1) & 1);
            tc_fence_after_sync();
            float u0[2][8];
#pragma unroll
            for (int tt = 0; tt < 2; ++tt) {
                const int t = half + 2 * tt;
                const int r = (t >> 1) * 16 + tyl, ox = (t & 1) * 8 + txl;
                float u[32];
                tmem_ld16(tlane + ACC3 + (uint32_t)b3 * 4 * N3 + (uint32_t)t * N3, u);
                tmem_ld16(tlane + ACC3 + (uint32_t)b3 * 4 * N3 + (uint32_t)t * N3 + 16, u + 16);
#pragma unroll
                for (int i = 0; i < 8; ++i) u0[tt][i] = u[i];
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    *reinterpret_cast<float4*>(sX1 + sw64_offset(r * 16 + ox, c)) = make_float4(u[8 + 4 * c], u[9 + 4 * c], u[10 + 4 * c], u[11 + 4 * c]);
                    *reinterpret_cast<float4*>(sX2 + sw64_offset(r * 16 + ox, c)) = make_float4(u[16 + 4 * c], u[17 + 4 * c], u[18 + 4 * c], u[19 + 4 * c]);
                }
            }
            tc_fence_before_sync();
            mbar_arrive(&acc3_free[b3]);
            epi_sync();
#pragma unroll
            for (int tt = 0; tt < 2; ++tt) {
                const int t = half + 2 * tt;
                const int oy = (t >> 1) * 16 + tyl, ox = (t & 1) * 8 + txl;
                const int y = ty * OUT_Y + oy, x = tx * OUT_X + ox;
                if (oy < OUT_Y && ox < OUT_X && y < d.H && x < d.W) {
                    float z = P.b4;
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        const float4 a1 = *reinterpret_cast<const float4*>(sX1 + sw64_offset((oy + 1) * 16 + ox, c));
                        const float4 a2 = *reinterpret_cast<const float4*>(sX2 + sw64_offset((oy + 2) * 16 + ox, c));
                        const float2 lo = unpack2(fadd2(fadd2(fadd2(pack2(u0[tt][4 * c], u0[tt][4 * c + 1]), pack2(a1.x, a1.y)), pack2(a2.x, a2.y)),
                                                        pack2(P.shift3[4 * c], P.shift3[4 * c + 1])));
                        const float2 hi = unpack2(fadd2(fadd2(fadd2(pack2(u0[tt][4 * c + 2], u0[tt][4 * c + 3]), pack2(a1.z, a1.w)), pack2(a2.z, a2.w)),
                                                        pack2(P.shift3[4 * c + 2], P.shift3[4 * c + 3])));
                        z = fmaf(relu_round_tf32(lo.x), P.w4[4 * c], z);
                        z = fmaf(relu_round_tf32(lo.y), P.w4[4 * c + 1], z);
                        z = fmaf(relu_round_tf32(hi.x), P.w4[4 * c + 2], z);
                        z = fmaf(relu_round_tf32(hi.y), P.w4[4 * c + 3], z);
                    }
                    weight[((int64_t)m * d.H + y) * d.W + x] = 1.0f / (1.0f + expf(-z));
                }
            }
            epi_sync();                                    // exchange buffers free again
        };
        int it = 0, prev_item = -1;
        for (int item = blockIdx.x; item < d.nitems; item += gridDim.x, ++it) {
            const int tx = item % d.tiles_x, ty = (item / d.tiles_x) % d.tiles_y;
            // ---- layer 2: out2(y2, x2) = T0(y2) + T1(y2 + 1) + T2(y2 + 2) -> operand of layer 3
            mbar_wait(acc2_full, it & 1);
            tc_fence_after_sync();
            float t0[2][16];
#pragma unroll
            for (int tt = 0; tt < 2; ++tt) {
                const int t = half + 2 * tt;
                const int r = (t >> 1) * 16 + tyl, x2 = (t & 1) * 8 + txl;             // input row / pixel column of this lane
                const uint32_t ta = tlane + (uint32_t)t * N2;
                float t1[16], t2[16];
                tmem_ld16(ta, t0[tt]);
                tmem_ld16(ta + 16, t1);
                tmem_ld16(ta + 32, t2);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    *reinterpret_cast<float4*>(sX1 + sw64_offset(r * 16 + x2, c)) = make_float4(t1[4 * c], t1[4 * c + 1], t1[4 * c + 2], t1[4 * c + 3]);
                    *reinterpret_cast<float4*>(sX2 + sw64_offset(r * 16 + x2, c)) = make_float4(t2[4 * c], t2[4 * c + 1], t2[4 * c + 2], t2[4 * c + 3]);
                }
            }
            tc_fence_before_sync();
            mbar_arrive(acc2_free);                        // accumulators drained: layer 2 of the next tile may run
            if (it >= 1) mbar_wait(free2, (it - 1) & 1);   // layer 3 of the previous tile has read A2
            epi_sync();                                    // T[1], T[2] of every input row are in shared memory
#pragma unroll
            for (int tt = 0; tt < 2; ++tt) {
                const int t = half + 2 * tt;
                const int y2 = (t >> 1) * 16 + tyl, x2 = (t & 1) * 8 + txl;
                if (y2 >= AY - 2) continue;                // rows 30, 31 are no layer-2 outputs
                const int y = ty * OUT_Y - 1 + y2, x = tx * OUT_X - 1 + x2;             // A2 origin = tile origin - 1
                const bool inside = y >= 0 && y < d.H && x >= 0 && x < d.W;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const float4 a1 = *reinterpret_cast<const float4*>(sX1 + sw64_offset((y2 + 1) * 16 + x2, c));
                    const float4 a2 = *reinterpret_cast<const float4*>(sX2 + sw64_offset((y2 + 2) * 16 + x2, c));
                    // ((T0 + T1) + T2) + shift, two channels per packed add
                    const float2 lo = unpack2(fadd2(fadd2(fadd2(pack2(t0[tt][4 * c], t0[tt][4 * c + 1]), pack2(a1.x, a1.y)), pack2(a2.x, a2.y)),
                                                    pack2(P.shift2[4 * c], P.shift2[4 * c + 1])));
                    const float2 hi = unpack2(fadd2(fadd2(fadd2(pack2(t0[tt][4 * c + 2], t0[tt][4 * c + 3]), pack2(a1.z, a1.w)), pack2(a2.z, a2.w)),
                                                    pack2(P.shift2[4 * c + 2], P.shift2[4 * c + 3])));
                    float4 v;
                    v.x = inside ? relu_round_tf32(lo.x) : 0.f;
                    v.y = inside ? relu_round_tf32(lo.y) : 0.f;
                    v.z = inside ? relu_round_tf32(hi.x) : 0.f;
                    v.w = inside ? relu_round_tf32(hi.y) : 0.f;
                    *reinterpret_cast<float4*>(sA2 + sw64_offset(y2 * AX + x2, c)) = v;
                }
            }
            fence_proxy_async_smem();
            mbar_arrive(full2);                            // -> the tensor core starts layer 3 of this tile ...
            epi_sync();                                    // exchange buffers free again
            if (prev_item >= 0) finish(it - 1, prev_item); // ... while the previous tile's layer-3 result is finished here
            prev_item = item;
        }
        if (prev_item >= 0) finish(it - 1, prev_item);
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

}  // namespace visf
}  // namespace tc
}  // namespace mvs

// entropy [M,H,W] -> visibility weight [M,H,W].  params_host: w1[16][9] b1[16] shift2[16] shift3[8] w4[8] b4 (BN folded);
// w2 / w3: device, TF32, packed [kw][Cin/4][kh][n][4] with n = 16 (w2) and n = 8 + 8 rows of zero padding after the three
// kh blocks (w3: 32 rows per (kw, quad)) — mvsformer_b200.engine.pack_vis_fused_weights.
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
