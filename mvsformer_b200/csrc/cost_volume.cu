// cost_volume.cu — A3/A4/A6: fused homography warp + group-wise correlation + visibility-weighted
// view aggregation (models/mvsformer_model.py:61-105) without ever materialising the
// N x C x D x H x W warped tensor.
//
// The visibility weight of a source view depends on the entropy of that view's whole depth
// column through a 7x7-receptive-field CNN (:87-91), so aggregation needs two sampling passes:
//   pass A  (cv_entropy_kernel)   per pixel and source view: s[d] = sum_g corr[g,d] -> softmax ->
//                                 entropy; fused with the eval-only cosine similarity (:81-85).
//   (vis net, vis_net.cu)         entropy -> weight
//   pass B  (cv_aggregate_kernel) per (pixel, d): volume[g] = sum_v w_v corr_v[g] / (sum_v w_v + 1e-6),
//                                 written channels-last [B,D,H,W,G] for the 3D CNN.
// Features are read in the reference's NCHW layout: lanes run along x, so each per-channel tap
// load of a warp touches one or two 128-byte lines of a channel plane.
//
// Roofline: algorithmic HBM bytes per stage = 4 [(N+1) C h w + D h w + G D h w];
// the sampling itself is bounded by L1/LSU throughput (4 taps x C x 4 B per (view, d, pixel)).
#include <float.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace mvs {

struct CvParams {
    const float* features;  // view 0 of batch 0
    int64_t batch_stride, view_stride;
    const float* relproj;   // [B, N, 12]
    const float* depth;     // [B, D, H, W]
    int N, C, G, D, H, W;
};

// ------------------------------------------------------------------------------------------
// pass A: one thread per pixel; the depth column lives in shared memory (lane-major).
// ------------------------------------------------------------------------------------------
template <int CPG_T, bool SIM>
__global__ void __launch_bounds__(128)
cv_entropy_kernel(CvParams p, float* __restrict__ entropy, float* __restrict__ sim_sum) {
    // per-thread depth columns in shared memory: s_col[k][tid] (scores of the current view) and,
    // for SIM, s_sim[k][tid] (cosine similarity summed over views); conflict-free (lane-major)
    extern __shared__ float s_dyn[];
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    float* s_col = s_dyn + tid;
    float* s_sim = s_dyn + (size_t)p.D * 128 + tid;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int b = blockIdx.z;
    if (x >= p.W || y >= p.H) return;
    const int cpg = CPG_T > 0 ? CPG_T : p.C / p.G;
    const int G = p.G;
    const int64_t hw = (int64_t)p.H * p.W;
    const int pix = y * p.W + x;
    const float* ref = p.features + (int64_t)b * p.batch_stride;
    const float half_w = (float)((p.W - 1) / 2.0), half_h = (float)((p.H - 1) / 2.0);
    const float inv_cpg = 1.0f / (float)cpg;

    if (SIM)
        for (int k = 0; k < p.D; ++k) s_sim[k * 128] = 0.0f;

    for (int v = 0; v < p.N; ++v) {
        const RelProj m = load_relproj(p.relproj + ((int64_t)b * p.N + v) * 12);
        const PixelRay ray = pixel_ray(m, (float)x, (float)y);
        const float* src = ref + (int64_t)(v + 1) * p.view_stride;
        float mx = -FLT_MAX;
#pragma unroll 1
        for (int k = 0; k < p.D; ++k) {
            const float dep = __ldg(p.depth + ((int64_t)b * p.D + k) * hw + pix);
            const Taps t = make_taps(m, ray, dep, p.H, p.W, half_w, half_h);
            float ssum = 0.0f, csum = 0.0f;
            for (int cp = 0; cp < cpg; ++cp) {
                float a = 0.0f, wn = 0.0f, rn = 0.0f;
                for (int g = 0; g < G; ++g) {
                    const int c = g * cpg + cp;
                    const float r = __ldg(ref + (int64_t)c * hw + pix);
                    const float wv = sample4(src + (int64_t)c * hw, t);
                    a = fmaf(r, wv, a);
                    if (SIM) {
                        wn = fmaf(wv, wv, wn);
                        rn = fmaf(r, r, rn);
                    }
                }
                ssum += a;
                if (SIM) {
                    // F.normalize(dim=group axis, eps=1e-12) on both operands (:82)
                    const float dn = fmaxf(sqrtf(rn), 1e-12f) * fmaxf(sqrtf(wn), 1e-12f);
                    csum += a / dn;
                }
            }
            const float sk = ssum * inv_cpg;
            s_col[k * 128] = sk;
            mx = fmaxf(mx, sk);
            if (SIM) s_sim[k * 128] += csum * inv_cpg;
        }
        // entropy of softmax over the depth column (:88-90)
        float den = 0.0f;
        for (int k = 0; k < p.D; ++k) {
            const float e = expf(s_col[k * 128] - mx);
            s_col[k * 128] = e;
            den += e;
        }
        float ent = 0.0f;
        for (int k = 0; k < p.D; ++k) {
            const float pr = s_col[k * 128] / den;
            ent -= pr * logf(pr + 1e-7f);
        }
        entropy[((int64_t)b * p.N + v) * hw + pix] = ent;
    }
    if (SIM)
        for (int k = 0; k < p.D; ++k) sim_sum[((int64_t)b * p.D + k) * hw + pix] = s_sim[k * 128];
}

// ------------------------------------------------------------------------------------------
// pass B: one thread per (pixel, depth); G accumulators; channels-last store.
// ------------------------------------------------------------------------------------------
template <int G_T, int CPG_T>
__global__ void __launch_bounds__(256)
cv_aggregate_kernel(CvParams p, const float* __restrict__ vis_weight, float* __restrict__ volume, int round_tf32) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int b = blockIdx.z / p.D, k = blockIdx.z % p.D;
    if (x >= p.W || y >= p.H) return;
    const int cpg = CPG_T > 0 ? CPG_T : p.C / G_T;
    const int64_t hw = (int64_t)p.H * p.W;
    const int pix = y * p.W + x;
    const float* ref = p.features + (int64_t)b * p.batch_stride;
    const float half_w = (float)((p.W - 1) / 2.0), half_h = (float)((p.H - 1) / 2.0);
    const float inv_cpg = 1.0f / (float)cpg;
    const float dep = __ldg(p.depth + ((int64_t)b * p.D + k) * hw + pix);

    float acc[G_T];
#pragma unroll
    for (int g = 0; g < G_T; ++g) acc[g] = 0.0f;
    float wsum = 0.0f;

    for (int v = 0; v < p.N; ++v) {
        const RelProj m = load_relproj(p.relproj + ((int64_t)b * p.N + v) * 12);
        const PixelRay ray = pixel_ray(m, (float)x, (float)y);
        const Taps t = make_taps(m, ray, dep, p.H, p.W, half_w, half_h);
        const float* src = ref + (int64_t)(v + 1) * p.view_stride;
        const float wv = __ldg(vis_weight + ((int64_t)b * p.N + v) * hw + pix);
        wsum += wv;
#pragma unroll
        for (int g = 0; g < G_T; ++g) {
            float a = 0.0f;
            for (int cp = 0; cp < cpg; ++cp) {
                const int c = g * cpg + cp;
                a = fmaf(__ldg(ref + (int64_t)c * hw + pix), sample4(src + (int64_t)c * hw, t), a);
            }
            acc[g] = fmaf(a * inv_cpg, wv, acc[g]);   // :101  volume_sum += in_prod_vol * vis_weight
        }
    }
    const float inv = 1.0f / (wsum + 1e-6f);          // :105
#pragma unroll
    for (int g = 0; g < G_T; ++g) {
        acc[g] *= inv;
        if (round_tf32) acc[g] = round_to_tf32(acc[g]);
    }
    float* out = volume + ((((int64_t)b * p.D + k) * p.H + y) * p.W + x) * G_T;
    if (G_T % 4 == 0) {
#pragma unroll
        for (int g = 0; g < G_T; g += 4)
            *reinterpret_cast<float4*>(out + g) = make_float4(acc[g], acc[g + 1], acc[g + 2], acc[g + 3]);
    } else {
#pragma unroll
        for (int g = 0; g < G_T; ++g) out[g] = acc[g];
    }
}

__global__ void argmax_gather_kernel(const float* __restrict__ score, const float* __restrict__ depth,
                                     float* __restrict__ out, int D, int64_t hw, int64_t total) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int64_t b = i / hw, pix = i % hw;
    float best = -FLT_MAX;
    int arg = 0;
    bool seen_nan = false;
    for (int k = 0; k < D; ++k) {
        const float s = __ldg(score + (b * D + k) * hw + pix);
        // torch.argmax: first maximal element; a NaN is maximal
        if (!seen_nan && (s != s)) { arg = k; seen_nan = true; }
        if (!seen_nan && s > best) { best = s; arg = k; }
    }
    out[i] = __ldg(depth + (b * D + arg) * hw + pix);
}

template <bool SIM>
static int dispatch_entropy(const CvParams& p, float* entropy, float* sim_sum, int B, cudaStream_t st) {
    dim3 block(32, 4);
    dim3 grid(cdiv(p.W, 32), cdiv(p.H, 4), B);
    const size_t smem = (size_t)p.D * 128 * sizeof(float) * (SIM ? 2 : 1);
    const int cpg = p.C / p.G;
#define MVS_LAUNCH_ENT(CPG)                                                                                   \
    {                                                                                                         \
        auto kern = cv_entropy_kernel<CPG, SIM>;                                                              \
        MVS_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));      \
        kern<<<grid, block, smem, st>>>(p, entropy, sim_sum);                                                 \
    }
    switch (cpg) {
        case 1: MVS_LAUNCH_ENT(1) break;
        case 2: MVS_LAUNCH_ENT(2) break;
        case 4: MVS_LAUNCH_ENT(4) break;
        case 8: MVS_LAUNCH_ENT(8) break;
        default: MVS_LAUNCH_ENT(0) break;
    }
#undef MVS_LAUNCH_ENT
    MVS_LAUNCH_OK("cv_entropy_kernel");
    return MVS_OK;
}

static int check_cv_args(const char* fn, const float* features, const float* relproj, const float* depth, int B, int V,
                         int C, int G, int D, int H, int W) {
    MVS_REQUIRE(features && relproj && depth, "%s: null pointer", fn);
    MVS_REQUIRE(B >= 1 && V >= 2 && C >= 1 && G >= 1 && D >= 1 && H >= 1 && W >= 1,
                "%s: empty shape B=%d V=%d C=%d G=%d D=%d H=%d W=%d", fn, B, V, C, G, D, H, W);
    MVS_REQUIRE(V - 1 <= MVS_MAX_SRC_VIEWS, "%s: at most %d source views (got %d)", fn, MVS_MAX_SRC_VIEWS, V - 1);
    MVS_REQUIRE(C % G == 0, "%s: channels (%d) not divisible by groups (%d)", fn, C, G);
    MVS_REQUIRE(D <= 64, "%s: at most 64 depth hypotheses per stage (got %d)", fn, D);
    MVS_REQUIRE((int64_t)B * D <= 65535, "%s: B*D = %lld exceeds 65535", fn, (long long)B * D);
    return MVS_OK;
}

// ------------------------------------------------------------------------------------------
// Streaming aggregation over STORED per-view correlations (opt-in, MVS_CV_STORE=1): the second
// sampling pass is replaced by one read of corr [B,N,D,H,W,8] written by the STORE variant of
// pass A.  Same FMA sequence as pass B (acc = fma(corr_v, w_v, acc) in view order, then * 1/(sum w + 1e-6)),
// so the volume is bit-identical.  One thread per voxel: N x 32 B in, 32 B out — a pure HBM stream.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
corr_aggregate_kernel(const float* __restrict__ corr, const float* __restrict__ vis_weight, float* __restrict__ volume,
                      int N, int D, int64_t hw, int64_t total, int round_tf32) {
    const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;          // (b, k, pixel)
    if (idx >= total) return;
    const int64_t pix = idx % hw;
    const int64_t bk = idx / hw;
    const int k = (int)(bk % D);
    const int64_t b = bk / D;
    float acc[8];
#pragma unroll
    for (int g = 0; g < 8; ++g) acc[g] = 0.0f;
    float wsum = 0.0f;
    for (int v = 0; v < N; ++v) {
        const float wv = __ldg(vis_weight + (b * N + v) * hw + pix);
        wsum += wv;
        const float4* c = reinterpret_cast<const float4*>(corr + ((((b * N + v) * D + k) * hw + pix) << 3));
        const float4 c0 = __ldg(c), c1 = __ldg(c + 1);
        acc[0] = fmaf(c0.x, wv, acc[0]); acc[1] = fmaf(c0.y, wv, acc[1]);
        acc[2] = fmaf(c0.z, wv, acc[2]); acc[3] = fmaf(c0.w, wv, acc[3]);
        acc[4] = fmaf(c1.x, wv, acc[4]); acc[5] = fmaf(c1.y, wv, acc[5]);
        acc[6] = fmaf(c1.z, wv, acc[6]); acc[7] = fmaf(c1.w, wv, acc[7]);
    }
    const float inv = 1.0f / (wsum + 1e-6f);
#pragma unroll
    for (int g = 0; g < 8; ++g) {
        acc[g] *= inv;
        if (round_tf32) acc[g] = round_to_tf32(acc[g]);
    }
    float4* out = reinterpret_cast<float4*>(volume + (idx << 3));
    out[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
    out[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
}

}  // namespace mvs

extern "C" int mvs_cost_volume_entropy(const float* features, int64_t batch_stride, int64_t view_stride,
                                       const float* relproj, const float* depth, float* entropy, float* sim_sum,
                                       int B, int V, int C, int G, int D, int H, int W, void* stream) {
    int rc = mvs::check_cv_args("mvs_cost_volume_entropy", features, relproj, depth, B, V, C, G, D, H, W);
    if (rc) return rc;
    MVS_REQUIRE(entropy, "mvs_cost_volume_entropy: null entropy output");
    mvs::CvParams p{features, batch_stride, view_stride, relproj, depth, V - 1, C, G, D, H, W};
    cudaStream_t st = (cudaStream_t)stream;
    return sim_sum ? mvs::dispatch_entropy<true>(p, entropy, sim_sum, B, st)
                   : mvs::dispatch_entropy<false>(p, entropy, nullptr, B, st);
}

extern "C" int mvs_corr_aggregate(const float* corr, const float* vis_weight, float* volume, int B, int N, int D, int H,
                                  int W, int round_tf32, void* stream) {
    MVS_REQUIRE(corr && vis_weight && volume, "mvs_corr_aggregate: null pointer");
    MVS_REQUIRE(B >= 1 && N >= 1 && D >= 1 && H >= 1 && W >= 1, "mvs_corr_aggregate: empty shape");
    MVS_REQUIRE((((uintptr_t)corr | (uintptr_t)volume) & 15) == 0, "mvs_corr_aggregate: buffers must be 16-byte aligned");
    const int64_t hw = (int64_t)H * W, total = hw * D * B;
    mvs::corr_aggregate_kernel<<<mvs::cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(corr, vis_weight, volume, N, D, hw,
                                                                                         total, round_tf32);
    MVS_LAUNCH_OK("corr_aggregate_kernel");
    return MVS_OK;
}

static int cost_volume_aggregate_impl(const float* features, int64_t batch_stride, int64_t view_stride,
                                      const float* relproj, const float* depth, const float* vis_weight,
                                      float* volume, int B, int V, int C, int G, int D, int H, int W, int round_tf32, void* stream) {
    int rc = mvs::check_cv_args("mvs_cost_volume_aggregate", features, relproj, depth, B, V, C, G, D, H, W);
    if (rc) return rc;
    MVS_REQUIRE(vis_weight && volume, "mvs_cost_volume_aggregate: null pointer");
    if (G != 8) MVS_UNSUPPORTED("mvs_cost_volume_aggregate: only G = 8 groups is built (got %d)", G);
    mvs::CvParams p{features, batch_stride, view_stride, relproj, depth, V - 1, C, G, D, H, W};
    cudaStream_t st = (cudaStream_t)stream;
    dim3 block(32, 8);
    dim3 grid(mvs::cdiv(W, 32), mvs::cdiv(H, 8), B * D);
    switch (C / G) {
        case 1: mvs::cv_aggregate_kernel<8, 1><<<grid, block, 0, st>>>(p, vis_weight, volume, round_tf32); break;
        case 2: mvs::cv_aggregate_kernel<8, 2><<<grid, block, 0, st>>>(p, vis_weight, volume, round_tf32); break;
        case 4: mvs::cv_aggregate_kernel<8, 4><<<grid, block, 0, st>>>(p, vis_weight, volume, round_tf32); break;
        case 8: mvs::cv_aggregate_kernel<8, 8><<<grid, block, 0, st>>>(p, vis_weight, volume, round_tf32); break;
        default: mvs::cv_aggregate_kernel<8, 0><<<grid, block, 0, st>>>(p, vis_weight, volume, round_tf32); break;
    }
    MVS_LAUNCH_OK("cv_aggregate_kernel");
    return MVS_OK;
}

extern "C" int mvs_cost_volume_aggregate(const float* features, int64_t batch_stride, int64_t view_stride,
                                         const float* relproj, const float* depth, const float* vis_weight,
                                         float* volume, int B, int V, int C, int G, int D, int H, int W, void* stream) {
    return cost_volume_aggregate_impl(features, batch_stride, view_stride, relproj, depth, vis_weight, volume, B, V, C, G, D, H, W, 0, stream);
}

extern "C" int mvs_cost_volume_aggregate_tf32(const float* features, int64_t batch_stride, int64_t view_stride,
                                              const float* relproj, const float* depth, const float* vis_weight,
                                              float* volume, int B, int V, int C, int G, int D, int H, int W, void* stream) {
    return cost_volume_aggregate_impl(features, batch_stride, view_stride, relproj, depth, vis_weight, volume, B, V, C, G, D, H, W, 1, stream);
}

extern "C" int mvs_argmax_gather(const float* score, const float* depth, float* out, int B, int D, int H, int W,
                                 void* stream) {
    MVS_REQUIRE(score && depth && out, "mvs_argmax_gather: null pointer");
    MVS_REQUIRE(B >= 1 && D >= 1 && H >= 1 && W >= 1, "mvs_argmax_gather: empty shape");
    const int64_t hw = (int64_t)H * W, total = hw * B;
    mvs::argmax_gather_kernel<<<mvs::cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(score, depth, out, D, hw, total);
    MVS_LAUNCH_OK("argmax_gather_kernel");
    return MVS_OK;
}
