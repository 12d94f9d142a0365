// cost_volume_tma.cu — the production cost-volume kernels: fused homography warp + group-wise
// correlation + entropy / visibility-weighted aggregation (models/mvsformer_model.py:61-105) with
// TMA-staged source-feature tiles in shared memory.
//
// A CTA owns a 32 x TH pixel tile of the reference view (256 threads = pixels x depth groups).
// For every source view it
//   1. projects its (pixel, hypothesis) samples, block-reduces their bounding box in the source
//      image (exact: every sample takes part, no monotonicity assumption),
//   2. pulls the box  [channels-of-chunk][BH][BW]  of the source feature map into shared memory
//      with ONE cp.async.bulk.tensor (TMA, 5-D tile over (x, y, c', g, view); out-of-image parts
//      are zero-filled by the TMA unit = grid_sample's zero padding),
//   3. bilinearly samples from shared memory: 4 LDS with immediate offsets per channel, weights
//      and the box offset cached per hypothesis.  A sample whose 2x2 footprint is not inside the
//      box (wild geometry) takes a predicated global-memory path with the reference's per-tap
//      bounds checks, so results never depend on the box heuristic.
// The N x C x D x H x W warped tensor, the sampling grid and the per-view correlation volumes
// never exist in HBM: pass A writes one entropy map per view (+ the eval-only cosine-similarity
// volume), pass B writes the aggregated volume once, channels-last.
//
// Channel chunking keeps the tile within shared memory at stage 1 (C = 64): pass A chunks over
// c' (all 8 groups of a few c'), because the cosine similarity normalises across groups;
// pass B chunks over groups.  Both are boxes of the same 5-D tensor map.
#include <cuda.h>
#include <float.h>
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"
#include "tcgen05.cuh"

namespace mvs {
namespace k1 {

using tc::mbar_init;
using tc::mbar_fence_init;
using tc::mbar_wait;
using tc::smem_u32;

struct K1Params {
    const float* feat;       // view 0 of batch 0, dense [B, V, C, H, W]
    const float* relproj;    // [B, N, 12]
    const float* depth;      // [B, D, H, W]
    int N, V, H, W;
    float* entropy;          // pass A out [B, N, H, W]
    float* sim_sum;          // pass A out [B, D, H, W] (SIM)
    float* corr;             // pass A out [B, N, D, H, W, 8] (STORE): per-view group correlation
    const float* vis_weight; // pass B in  [B, N, H, W]
    float* volume;           // pass B out [B, D, H, W, 8]
    int round_tf32;          // pass B: round the stored volume to TF32
};

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3,
                                            int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}

// a / b with a correctly rounded reciprocal rb = RN(1/b): q = RN(a*rb); r = a - b*q (exact, one FMA);
// RN(q + r*rb).  This is the refinement IEEE division itself uses (Markstein), so the quotient is the
// correctly rounded one — bit-identical to __fdiv_rn for operands in the normal range (checked
// exhaustively on random operands against exact rational arithmetic) — but the reciprocal is shared:
// one for both projected coordinates, one per kernel for the image-size constants.  Samples whose
// coordinates leave the normal range fail the `sane` test and take the predicated global path, which
// uses plain IEEE divisions (make_taps).
__device__ __forceinline__ float div_by(float a, float b, float rb) {
    const float q = __fmul_rn(a, rb);
    const float r = __fmaf_rn(-b, q, a);
    return __fmaf_rn(r, rb, q);
}

// Sample position of (pixel ray, depth): the arithmetic of make_taps() in common.cuh.
__device__ __forceinline__ void project(const RelProj& m, const PixelRay& ray, float depth, int H, int W, float half_w,
                                        float half_h, float rcp_half_w, float rcp_half_h, float* ix, float* iy) {
    const float qx = __fadd_rn(__fmul_rn(ray.x, depth), m.t0);
    const float qy = __fadd_rn(__fmul_rn(ray.y, depth), m.t1);
    const float qz = __fadd_rn(__fmul_rn(ray.z, depth), m.t2);
    const float den = __fadd_rn(qz, 1e-6f);
    const float rden = __frcp_rn(den);
    const float gx = __fadd_rn(div_by(div_by(qx, den, rden), half_w, rcp_half_w), -1.0f);
    const float gy = __fadd_rn(div_by(div_by(qy, den, rden), half_h, rcp_half_h), -1.0f);
    *ix = __fmul_rn(__fmul_rn(__fadd_rn(gx, 1.0f), 0.5f), (float)(W - 1));
    *iy = __fmul_rn(__fmul_rn(__fadd_rn(gy, 1.0f), 0.5f), (float)(H - 1));
}

// MINB = CTAs per SM the register allocation must allow (co-resident CTAs hide each other's TMA waits).
// STORE (pass A only, opt-in, MVS_CV_STORE=1): also write every view's group correlation, so that the
// aggregation becomes one streaming pass over it (corr_aggregate_kernel, cost_volume.cu) instead of a second
// sampling pass.  Only instantiated where the correlation is smaller than the warped tensor (C/G >= 2).
template <int CPG, int DG, int KPT, int BW, int BH, int NCH, bool PASS_B, bool SIM, bool STORE = false,
          int MINB = (CPG == 1 ? (PASS_B ? 3 : 4) : 2)>
__global__ void __launch_bounds__(256, MINB)
cost_volume_kernel(const __grid_constant__ CUtensorMap tmap, K1Params p) {
    constexpr int G = 8, C = G * CPG, D = DG * KPT;
    constexpr int TP = 256 / DG, TH = 8 / DG;                    // pixels per CTA, tile height (width 32)
    constexpr int CC = C / NCH;                                  // channels per chunk
    constexpr int CPC = PASS_B ? CPG : CPG / NCH;                // c' per chunk
    constexpr int GPC = PASS_B ? G / NCH : G;                    // groups per chunk
    constexpr int PLANE = BH * BW;
    static_assert(CPC * GPC == CC && CPC >= 1 && GPC >= 1, "chunking");

    extern __shared__ __align__(128) uint8_t smem[];
    float* tile = reinterpret_cast<float*>(smem);                // [CC][BH][BW]
    float* s_ref = tile + CC * PLANE;                            // [C][TP]
    float* s_col = s_ref + C * TP;                               // [D][TP] (pass A)
    float* s_red = s_col + (PASS_B ? 0 : D * TP);                // [8][4]
    int* s_org = reinterpret_cast<int*>(s_red + 32);             // bx0, by0
    uint64_t* bar = reinterpret_cast<uint64_t*>(s_org + 2);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int dg = tid / TP, pix = tid - dg * TP;
    const int x = blockIdx.x * 32 + (pix & 31), y = blockIdx.y * TH + (pix >> 5);
    const int b = blockIdx.z;
    const bool live = x < p.W && y < p.H;
    const int64_t hw = (int64_t)p.H * p.W;
    const int pixoff = live ? y * p.W + x : 0;
    const float* ref = p.feat + (int64_t)b * p.V * C * hw;
    const float half_w = (float)((p.W - 1) / 2.0), half_h = (float)((p.H - 1) / 2.0);
    const float rcp_half_w = __frcp_rn(half_w), rcp_half_h = __frcp_rn(half_h);
    const float inv_cpg = 1.0f / (float)CPG;

    if (tid == 0) { mbar_init(bar, 1); mbar_fence_init(); }
    for (int i = tid; i < C * TP; i += 256) {
        const int c = i / TP, q = i - c * TP;
        const int qx = blockIdx.x * 32 + (q & 31), qy = blockIdx.y * TH + (q >> 5);
        s_ref[i] = (qx < p.W && qy < p.H) ? __ldg(ref + (int64_t)c * hw + (int64_t)qy * p.W + qx) : 0.0f;
    }
    __syncthreads();

    float rden[CPG];                                             // F.normalize denominators of the reference (SIM)
    if (SIM && !PASS_B) {
#pragma unroll
        for (int cp = 0; cp < CPG; ++cp) {
            float rn = 0.0f;
#pragma unroll
            for (int g = 0; g < G; ++g) { const float r = s_ref[(g * CPG + cp) * TP + pix]; rn = fmaf(r, r, rn); }
            rden[cp] = fmaxf(sqrtf(rn), 1e-12f);
        }
    }

    float acc[PASS_B ? G : 1][KPT];                              // pass B: weighted correlation sums
    float cosv[KPT];                                             // pass A: cosine similarity summed over views
#pragma unroll
    for (int j = 0; j < KPT; ++j) {
        cosv[j] = 0.0f;
#pragma unroll
        for (int g = 0; g < (PASS_B ? G : 1); ++g) acc[g][j] = 0.0f;
    }
    float wsum = 0.0f;
    uint32_t phase = 0;

    for (int v = 0; v < p.N; ++v) {
        const RelProj m = load_relproj(p.relproj + ((int64_t)b * p.N + v) * 12);
        const PixelRay ray = pixel_ray(m, (float)x, (float)y);
        const float* src = ref + (int64_t)(v + 1) * C * hw;

        // ---- 1. sample positions of my hypotheses and the CTA's bounding box -------------------
        float ix[KPT], iy[KPT];
        float bminx = FLT_MAX, bminy = FLT_MAX;
#pragma unroll
        for (int j = 0; j < KPT; ++j) {
            const float dep = __ldg(p.depth + ((int64_t)b * D + dg * KPT + j) * hw + pixoff);
            project(m, ray, dep, p.H, p.W, half_w, half_h, rcp_half_w, rcp_half_h, &ix[j], &iy[j]);
            const bool sane = live && fabsf(ix[j]) < 1e7f && fabsf(iy[j]) < 1e7f;      // false for NaN / inf too
            if (sane) { bminx = fminf(bminx, ix[j]); bminy = fminf(bminy, iy[j]); }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            bminx = fminf(bminx, __shfl_xor_sync(0xffffffffu, bminx, o));
            bminy = fminf(bminy, __shfl_xor_sync(0xffffffffu, bminy, o));
        }
        if (lane == 0) { s_red[warp * 2] = bminx; s_red[warp * 2 + 1] = bminy; }
        __syncthreads();                                          // also: previous view's tile fully consumed
        if (tid == 0) {
            float mx = FLT_MAX, my = FLT_MAX;
            for (int w = 0; w < 8; ++w) { mx = fminf(mx, s_red[w * 2]); my = fminf(my, s_red[w * 2 + 1]); }
            // The TMA unit requires the innermost start coordinate to be 16-byte aligned (a multiple of 4
            // floats; negative values are fine) — measured on B200: any other value raises "illegal
            // instruction".  Round the box origin down; BW carries the 3 columns of slack.
            const int bx0 = mx < 1e7f ? ((int)floorf(mx) & ~3) : 0;
            const int by0 = my < 1e7f ? (int)floorf(my) : 0;
            s_org[0] = bx0; s_org[1] = by0;
            mbar_expect_tx(bar, CC * PLANE * 4);
            tma_load_5d(tile, &tmap, bar, bx0, by0, 0, 0, b * p.V + v + 1);
        }
        __syncthreads();
        const int bx0 = s_org[0], by0 = s_org[1];

        // ---- 2. per-hypothesis cache: box offset (or -1 = outside the box) and fractions ---------
        int off[KPT];
        float fx[KPT], fy[KPT];
#pragma unroll
        for (int j = 0; j < KPT; ++j) {
            const bool sane = live && fabsf(ix[j]) < 1e7f && fabsf(iy[j]) < 1e7f;
            const float x0 = floorf(ix[j]), y0 = floorf(iy[j]);
            fx[j] = ix[j] - x0; fy[j] = iy[j] - y0;
            const int lx = sane ? (int)x0 - bx0 : -1, ly = sane ? (int)y0 - by0 : -1;
            off[j] = (lx >= 0 && lx + 1 < BW && ly >= 0 && ly + 1 < BH) ? ly * BW + lx : -1;
        }
        float wv = 0.0f;
        if (PASS_B) {
            wv = live ? __ldg(p.vis_weight + ((int64_t)b * p.N + v) * hw + pixoff) : 0.0f;
            wsum += wv;
        }
        float sview[KPT];
        float ag[STORE ? G : 1][KPT];                            // STORE: this view's per-group sums, in pass B's FMA order
#pragma unroll
        for (int j = 0; j < KPT; ++j) {
            sview[j] = 0.0f;
#pragma unroll
            for (int g = 0; g < (STORE ? G : 1); ++g) ag[g][j] = 0.0f;
        }

        // ---- 3. channel chunks -------------------------------------------------------------------
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
            if (ch > 0) {
                __syncthreads();                                  // previous chunk consumed
                if (tid == 0) {
                    mbar_expect_tx(bar, CC * PLANE * 4);
                    tma_load_5d(tile, &tmap, bar, bx0, by0, PASS_B ? 0 : ch * CPC, PASS_B ? ch * GPC : 0, b * p.V + v + 1);
                }
            }
            mbar_wait(bar, phase);
            phase ^= 1u;
            const int cp0 = PASS_B ? 0 : ch * CPC, g0 = PASS_B ? ch * GPC : 0;
#pragma unroll
            for (int j = 0; j < KPT; ++j) {
                if (!live) continue;
                if (off[j] >= 0) {
                    const float* t = tile + off[j];
                    const float w00 = (1.0f - fx[j]) * (1.0f - fy[j]), w01 = fx[j] * (1.0f - fy[j]);
                    const float w10 = (1.0f - fx[j]) * fy[j], w11 = fx[j] * fy[j];
                    if (PASS_B) {
#pragma unroll
                        for (int gl = 0; gl < GPC; ++gl) {
                            float a = 0.0f;
#pragma unroll
                            for (int cp = 0; cp < CPG; ++cp) {
                                const float* q = t + (gl * CPG + cp) * PLANE;
                                float s = q[0] * w00;
                                s = fmaf(q[1], w01, s); s = fmaf(q[BW], w10, s); s = fmaf(q[BW + 1], w11, s);
                                a = fmaf(s_ref[((g0 + gl) * CPG + cp) * TP + pix], s, a);
                            }
                            acc[PASS_B ? g0 + gl : 0][j] = fmaf(a * inv_cpg, wv, acc[PASS_B ? g0 + gl : 0][j]);
                        }
                    } else {
#pragma unroll
                        for (int cpl = 0; cpl < CPC; ++cpl) {
                            float a = 0.0f, wn = 0.0f;
#pragma unroll
                            for (int g = 0; g < G; ++g) {
                                const float* q = t + (g * CPC + cpl) * PLANE;
                                float s = q[0] * w00;
                                s = fmaf(q[1], w01, s); s = fmaf(q[BW], w10, s); s = fmaf(q[BW + 1], w11, s);
                                const float r = s_ref[(g * CPG + cp0 + cpl) * TP + pix];
                                a = fmaf(r, s, a);
                                if (STORE) ag[STORE ? g : 0][j] = fmaf(r, s, ag[STORE ? g : 0][j]);
                                if (SIM) wn = fmaf(s, s, wn);
                            }
                            sview[j] += a;
                            if (SIM) cosv[j] += a / (rden[SIM ? cp0 + cpl : 0] * fmaxf(sqrtf(wn), 1e-12f));
                        }
                    }
                } else {
                    // predicated global path: the reference's per-tap bounds checks, any geometry
                    const float dep = __ldg(p.depth + ((int64_t)b * D + dg * KPT + j) * hw + pixoff);
                    const Taps tp = make_taps(m, ray, dep, p.H, p.W, half_w, half_h);
                    if (PASS_B) {
                        for (int gl = 0; gl < GPC; ++gl) {
                            float a = 0.0f;
                            for (int cp = 0; cp < CPG; ++cp) {
                                const int c = (g0 + gl) * CPG + cp;
                                a = fmaf(s_ref[c * TP + pix], sample4(src + (int64_t)c * hw, tp), a);
                            }
#pragma unroll
                            for (int gg = 0; gg < G; ++gg)
                                if (gg == g0 + gl) acc[PASS_B ? gg : 0][j] = fmaf(a * inv_cpg, wv, acc[PASS_B ? gg : 0][j]);
                        }
                    } else {
                        for (int cpl = 0; cpl < CPC; ++cpl) {
                            float a = 0.0f, wn = 0.0f;
                            for (int g = 0; g < G; ++g) {
                                const int c = g * CPG + cp0 + cpl;
                                const float s = sample4(src + (int64_t)c * hw, tp);
                                const float r = s_ref[c * TP + pix];
                                a = fmaf(r, s, a);
                                if (STORE) {
#pragma unroll
                                    for (int gg = 0; gg < G; ++gg)
                                        if (gg == g) ag[STORE ? gg : 0][j] = fmaf(r, s, ag[STORE ? gg : 0][j]);
                                }
                                if (SIM) wn = fmaf(s, s, wn);
                            }
                            sview[j] += a;
                            if (SIM) {
                                float rd = 0.0f;
#pragma unroll
                                for (int cc = 0; cc < CPG; ++cc)
                                    if (cc == cp0 + cpl) rd = rden[SIM ? cc : 0];
                                cosv[j] += a / (rd * fmaxf(sqrtf(wn), 1e-12f));
                            }
                        }
                    }
                }
            }
        }

        if (STORE && live) {
#pragma unroll
            for (int j = 0; j < KPT; ++j) {
                float* out = p.corr + ((((int64_t)b * p.N + v) * D + dg * KPT + j) * hw + pixoff) * G;
                *reinterpret_cast<float4*>(out) = make_float4(ag[0][j] * inv_cpg, ag[STORE ? 1 : 0][j] * inv_cpg,
                                                              ag[STORE ? 2 : 0][j] * inv_cpg, ag[STORE ? 3 : 0][j] * inv_cpg);
                *reinterpret_cast<float4*>(out + 4) = make_float4(ag[STORE ? 4 : 0][j] * inv_cpg, ag[STORE ? 5 : 0][j] * inv_cpg,
                                                                  ag[STORE ? 6 : 0][j] * inv_cpg, ag[STORE ? 7 : 0][j] * inv_cpg);
            }
        }

        // ---- 4. pass A: entropy of softmax over the full depth column ---------------------------------
        if (!PASS_B) {
#pragma unroll
            for (int j = 0; j < KPT; ++j) s_col[(dg * KPT + j) * TP + pix] = sview[j] * inv_cpg;
            __syncthreads();
            if (dg == 0 && live) {
                float mx = -FLT_MAX;
                for (int k = 0; k < D; ++k) mx = fmaxf(mx, s_col[k * TP + pix]);
                float den = 0.0f;
                for (int k = 0; k < D; ++k) den += expf(s_col[k * TP + pix] - mx);
                float ent = 0.0f;
                for (int k = 0; k < D; ++k) {
                    const float pr = expf(s_col[k * TP + pix] - mx) / den;
                    ent -= pr * logf(pr + 1e-7f);
                }
                p.entropy[((int64_t)b * p.N + v) * hw + pixoff] = ent;
            }
        }
    }

    if (!live) return;
    if (PASS_B) {
        const float inv = 1.0f / (wsum + 1e-6f);
#pragma unroll
        for (int j = 0; j < KPT; ++j) {
#pragma unroll
            for (int g = 0; g < (PASS_B ? G : 1); ++g) {
                acc[g][j] *= inv;
                if (p.round_tf32) acc[g][j] = tc::to_tf32(acc[g][j]);
            }
            float* out = p.volume + (((int64_t)b * D + dg * KPT + j) * hw + pixoff) * G;
            *reinterpret_cast<float4*>(out) = make_float4(acc[0][j], acc[PASS_B ? 1 : 0][j], acc[PASS_B ? 2 : 0][j], acc[PASS_B ? 3 : 0][j]);
            *reinterpret_cast<float4*>(out + 4) = make_float4(acc[PASS_B ? 4 : 0][j], acc[PASS_B ? 5 : 0][j], acc[PASS_B ? 6 : 0][j], acc[PASS_B ? 7 : 0][j]);
        }
    } else if (SIM) {
#pragma unroll
        for (int j = 0; j < KPT; ++j) p.sim_sum[((int64_t)b * D + dg * KPT + j) * hw + pixoff] = cosv[j] * inv_cpg;
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

// 5-D map over the dense feature tensor [B*V][G][CPG][H][W] (innermost first: x, y, c', g, view).
static int make_feature_map(CUtensorMap* map, const float* feat, int BV, int G, int CPG, int H, int W, int bw, int bh, int box_cp,
                            int box_g) {
    EncodeTiledFn fn = encode_fn();
    MVS_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
    const cuuint64_t dims[5] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)CPG, (cuuint64_t)G, (cuuint64_t)BV};
    const cuuint64_t hw = (cuuint64_t)H * W;
    const cuuint64_t strides[4] = {(cuuint64_t)W * 4, hw * 4, hw * 4 * CPG, hw * 4 * CPG * G};
    const cuuint32_t box[5] = {(cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)box_cp, (cuuint32_t)box_g, 1};
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult rc = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(feat), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MVS_REQUIRE(rc == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with code %d", (int)rc);
    return MVS_OK;
}

template <int CPG, int DG, int KPT, int BW, int BH, int NCH, bool PASS_B, bool SIM, bool STORE = false>
static int launch(const K1Params& p, int B, cudaStream_t st) {
    constexpr int G = 8, C = G * CPG, D = DG * KPT, TP = 256 / DG, TH = 8 / DG, CC = C / NCH;
    constexpr int CPC = PASS_B ? CPG : CPG / NCH, GPC = PASS_B ? G / NCH : G;
    const size_t smem = (size_t)(CC * BH * BW + C * TP + (PASS_B ? 0 : D * TP) + 32 + 2) * 4 + 16;
    CUtensorMap map;
    int rc = make_feature_map(&map, p.feat, B * p.V, G, CPG, p.H, p.W, BW, BH, CPC, GPC);
    if (rc) return rc;
    static_assert(!STORE || (!PASS_B && CPG >= 2), "STORE is a pass A variant for C/G >= 2");
    auto kern = cost_volume_kernel<CPG, DG, KPT, BW, BH, NCH, PASS_B, SIM, STORE>;
    MVS_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(cdiv(p.W, 32), cdiv(p.H, TH), B);
    kern<<<grid, 256, smem, st>>>(map, p);
    MVS_LAUNCH_OK("cost_volume_kernel");
    return MVS_OK;
}

// Box sizes BW x BH: the 32-pixel-wide tile + the sweep of the hypotheses along the epipolar line +
// the spread of the per-pixel hypotheses inside the tile + 3 columns of alignment slack.  They only
// steer how many samples take the fast path; any geometry stays correct (and bit-identical).  Measured
// on the bench workload (scripts/box_coverage.py, scripts/box_sweep.py): 128x10 / 64x8 / 64x12 / 80x16
// cover 99.7-100 % of the samples of every view; the first-round 112x8 / 48x12 / 48x16 boxes covered
// only 88-99 % at stages 3-4 and cost 0.39 ms per reference view in predicated global taps.  A row
// pitch that is a multiple of 32 floats also keeps image rows on the same banks (stage 1: 1.25 instead
// of 1.49 shared-memory wavefronts per load).
// Depth groups per stage (threads = 256 / DG pixels x DG groups): 8/4/2/1, i.e. four hypotheses per
// thread at every stage.  More groups = fewer registers per thread and more, smaller CTAs (better wave
// quantisation on 148 SMs); 4/2/1/2 groups were measured 5-20 % slower on B200 and are no longer built.
template <bool PASS_B, bool SIM>
static int dispatch(const K1Params& p, int B, int C, int D, cudaStream_t st) {
    if (C == 64 && D == 32) return launch<8, 8, 4, 128, 10, 4, PASS_B, SIM>(p, B, st);
    if (C == 32 && D == 16) return launch<4, 4, 4, 64, 8, 2, PASS_B, SIM>(p, B, st);
    if (C == 16 && D == 8) return launch<2, 2, 4, 64, 12, 1, PASS_B, SIM>(p, B, st);
    if (C == 8 && D == 4) return launch<1, 1, 4, 80, 16, 1, PASS_B, SIM>(p, B, st);
    return 1;   // not covered: caller uses the generic kernels
}

// pass A with the per-view correlation stored (stages with C/G >= 2 only; box sizes as above)
template <bool SIM>
static int dispatch_store(const K1Params& p, int B, int C, int D, cudaStream_t st) {
    if (C == 64 && D == 32) return launch<8, 8, 4, 128, 10, 4, false, SIM, true>(p, B, st);
    if (C == 32 && D == 16) return launch<4, 4, 4, 64, 8, 2, false, SIM, true>(p, B, st);
    if (C == 16 && D == 8) return launch<2, 2, 4, 64, 12, 1, false, SIM, true>(p, B, st);
    return 1;
}

}  // namespace k1

// Returns 1 when the shape is not covered by the TMA kernels (the generic kernels of
// cost_volume.cu then run), 0 on success, negative on error.
int cost_volume_tma_entropy(const float* features, int64_t batch_stride, int64_t view_stride, const float* relproj,
                            const float* depth, float* entropy, float* sim_sum, int B, int V, int C, int G, int D, int H, int W,
                            cudaStream_t st) {
    const int64_t hw = (int64_t)H * W;
    if (G != 8 || (W % 4) != 0 || view_stride != C * hw || batch_stride != V * C * hw || ((uintptr_t)features & 15)) return 1;
    k1::K1Params p{features, relproj, depth, V - 1, V, H, W, entropy, sim_sum, nullptr, nullptr, nullptr, 0};
    return sim_sum ? k1::dispatch<false, true>(p, B, C, D, st) : k1::dispatch<false, false>(p, B, C, D, st);
}

// Pass A that also stores corr [B,N,D,H,W,8]; returns 1 when the shape is not covered.
int cost_volume_tma_entropy_store(const float* features, int64_t batch_stride, int64_t view_stride, const float* relproj,
                                  const float* depth, float* entropy, float* sim_sum, float* corr, int B, int V, int C, int G,
                                  int D, int H, int W, cudaStream_t st) {
    const int64_t hw = (int64_t)H * W;
    if (G != 8 || (W % 4) != 0 || view_stride != C * hw || batch_stride != V * C * hw || ((uintptr_t)features & 15)) return 1;
    k1::K1Params p{features, relproj, depth, V - 1, V, H, W, entropy, sim_sum, corr, nullptr, nullptr, 0};
    return sim_sum ? k1::dispatch_store<true>(p, B, C, D, st) : k1::dispatch_store<false>(p, B, C, D, st);
}

int cost_volume_tma_aggregate(const float* features, int64_t batch_stride, int64_t view_stride, const float* relproj,
                              const float* depth, const float* vis_weight, float* volume, int B, int V, int C, int G, int D,
                              int H, int W, int round_tf32, cudaStream_t st) {
    const int64_t hw = (int64_t)H * W;
    if (G != 8 || (W % 4) != 0 || view_stride != C * hw || batch_stride != V * C * hw || ((uintptr_t)features & 15)) return 1;
    k1::K1Params p{features, relproj, depth, V - 1, V, H, W, nullptr, nullptr, nullptr, vis_weight, volume, round_tf32};
    return k1::dispatch<true, false>(p, B, C, D, st);
}

}  // namespace mvs
