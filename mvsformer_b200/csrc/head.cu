// head.cu — A7/A8/A9: the `prob` convolution, softmax / temperature-regression head, confidence
// and the depth-hypothesis schedulers.  All are single-pass HBM-bound kernels: one thread per
// pixel (the depth column, D <= 64, lives in registers) or per voxel, lanes along x.
//
//   mvs_prob_conv_cl          models/module.py:493 (3x3x3, no bias) / :582 (1x1x1 + bias)
//   mvs_regression_head       models/mvsformer_model.py:110-125, models/module.py:597-603
//   mvs_depth_regression      models/module.py:597-603
//   mvs_conf_regression       models/module.py:606-619
//   mvs_init_inverse_range    models/module.py:633-639      mvs_init_range      :622-630
//   mvs_schedule_inverse_range models/module.py:642-653     mvs_schedule_range  :687-699
//   mvs_confidence_accumulate models/mvsformer_model.py:438-442
#include <float.h>

#include "common.cuh"

namespace mvs {

constexpr int HEAD_DMAX = 64;

struct ProbWeights {
    float w[27][8];
    float bias;
};

// 1x1x1 `prob` conv + bias (CostRegNet3D / CostRegNet2D, models/module.py:582): one voxel per thread, two LDG.128.
__global__ void __launch_bounds__(256)
prob_conv1_kernel(const float* __restrict__ x, float* __restrict__ pre, int64_t total, const __grid_constant__ ProbWeights P) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const float4* p = reinterpret_cast<const float4*>(x + idx * 8);
    const float4 a = __ldg(p), c = __ldg(p + 1);
    float acc = P.bias;
    acc = fmaf(a.x, P.w[0][0], acc); acc = fmaf(a.y, P.w[0][1], acc); acc = fmaf(a.z, P.w[0][2], acc); acc = fmaf(a.w, P.w[0][3], acc);
    acc = fmaf(c.x, P.w[0][4], acc); acc = fmaf(c.y, P.w[0][5], acc); acc = fmaf(c.z, P.w[0][6], acc); acc = fmaf(c.w, P.w[0][7], acc);
    pre[idx] = acc;
}

// 3x3x3 `prob` conv (8 -> 1, CostRegNet.prob, models/module.py:491), marching form.  A one-voxel-per-thread kernel reads 54
// LDG.128 per voxel and was L1-bound (54 + 99 us at stages 1-2 for 85 MB of input, profiles/r02_launches_step.csv).
// Here a CTA owns a 32 x 8 pixel tile and marches over PZ output slices: each input slice tile (+1 halo, zero-filled
// outside the volume = the conv's zero padding) is staged ONCE in shared memory, split into its two channel halves so that
// a warp's LDS.128 is conflict-free, and feeds the three output slices it belongs to (partial sums a0/a1/a2 roll through
// registers).  The next slice is prefetched into registers under the FMAs; one __syncthreads per slice.
constexpr int PT_X = 32, PT_Y = 8, PH_X = PT_X + 2, PH_Y = PT_Y + 2, PZ = 8, PTILE = PH_X * PH_Y;

__device__ __forceinline__ float dot8(const float4& a, const float4& c, const float (&w)[8], float acc) {
    acc = fmaf(a.x, w[0], acc); acc = fmaf(a.y, w[1], acc); acc = fmaf(a.z, w[2], acc); acc = fmaf(a.w, w[3], acc);
    acc = fmaf(c.x, w[4], acc); acc = fmaf(c.y, w[5], acc); acc = fmaf(c.z, w[6], acc); acc = fmaf(c.w, w[7], acc);
    return acc;
}

__global__ void __launch_bounds__(256)
prob_conv3_tiled_kernel(const float* __restrict__ x, float* __restrict__ pre, int D, int H, int W, int zchunks,
                        const __grid_constant__ ProbWeights P) {
    __shared__ float4 s[2][2][PTILE];                       // [buffer][channel half][row * PH_X + col]
    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
    const int x0 = blockIdx.x * PT_X, y0 = blockIdx.y * PT_Y;
    const int64_t b = blockIdx.z / zchunks;
    const int z0 = (blockIdx.z % zchunks) * PZ, z1 = min(D, z0 + PZ);          // output slices [z0, z1)
    const int ox = x0 + tx, oy = y0 + ty;
    const bool live = ox < W && oy < H;

    float4 regs[3];
    auto fetch = [&](int iz) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int e = tid + 256 * k;                     // (voxel, half) pairs of the slice tile, 16-byte granules
            const int v = e >> 1, row = v / PH_X, col = v - row * PH_X;
            const int gy = y0 - 1 + row, gx = x0 - 1 + col;
            float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
            if (e < 2 * PTILE && iz >= 0 && iz < D && gy >= 0 && gy < H && gx >= 0 && gx < W)
                r = __ldg(reinterpret_cast<const float4*>(x + (((b * D + iz) * H + gy) * (int64_t)W + gx) * 8) + (e & 1));
            regs[k] = r;
        }
    };
    auto stash = [&](int buf) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int e = tid + 256 * k;
            if (e < 2 * PTILE) s[buf][e & 1][e >> 1] = regs[k];
        }
    };

    fetch(z0 - 1);
    stash(0);
    __syncthreads();
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;                      // partial sums of out[zi-1], out[zi], out[zi+1]
    for (int zi = z0 - 1; zi <= z1; ++zi) {
        const int buf = (zi - (z0 - 1)) & 1;
        if (zi < z1) fetch(zi + 1);                          // in flight under the FMAs below
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int v = (ty + ky) * PH_X + tx + kx;
                const float4 a = s[buf][0][v], c = s[buf][1][v];
                a2 = dot8(a, c, P.w[ky * 3 + kx], a2);               // kz = 0: this slice is the one below out[zi+1]
                a1 = dot8(a, c, P.w[9 + ky * 3 + kx], a1);           // kz = 1
                a0 = dot8(a, c, P.w[18 + ky * 3 + kx], a0);          // kz = 2
            }
        }
        const int oz = zi - 1;
        if (live && oz >= z0 && oz < z1) pre[((b * D + oz) * H + oy) * (int64_t)W + ox] = a0 + P.bias;
        a0 = a1; a1 = a2; a2 = 0.f;
        if (zi < z1) stash(buf ^ 1);
        __syncthreads();
    }
}

// One thread per pixel.  D <= HEAD_DMAX; the column is re-read from L1/L2 for the second softmax
// instead of being held in registers (keeps the kernel generic in D).
__global__ void __launch_bounds__(256)
regression_head_kernel(const float* __restrict__ pre, const float* __restrict__ dv, float tmp, int mode,
                       float* __restrict__ prob, float* __restrict__ depth, float* __restrict__ conf, int D, int64_t hw) {
    const int64_t pix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;       // grid.y = batch item: no 64-bit division
    if (pix >= hw) return;
    const int64_t b = blockIdx.y, i = b * hw + pix;
    const float* col = pre + b * D * hw + pix;
    const float* dcol = dv + b * D * hw + pix;
    float mx = -FLT_MAX;
    int arg = 0;
    for (int k = 0; k < D; ++k) {
        const float v = __ldg(col + k * hw);
        if (v > mx) { mx = v; arg = k; }
    }
    float den = 0.0f;
    for (int k = 0; k < D; ++k) den += expf(__ldg(col + k * hw) - mx);
    if (prob) {
        float* pcol = prob + b * D * hw + pix;
        for (int k = 0; k < D; ++k) pcol[k * hw] = expf(__ldg(col + k * hw) - mx) / den;
    }
    conf[i] = 1.0f / den;                       // max_d softmax = exp(0) / den   (:125)
    if (mode == 1) {
        depth[i] = __ldg(dcol + arg * hw);      // :117-120 (first maximal index, as torch.max)
    } else {
        // softmax(pre * tmp): the maximum of the scaled column is tmp*mx for tmp >= 0, tmp*min otherwise
        float smx = -FLT_MAX;
        for (int k = 0; k < D; ++k) smx = fmaxf(smx, __ldg(col + k * hw) * tmp);
        float sden = 0.0f, num = 0.0f;
        for (int k = 0; k < D; ++k) {
            const float e = expf(__ldg(col + k * hw) * tmp - smx);
            sden += e;
            num = fmaf(e, __ldg(dcol + k * hw), num);
        }
        depth[i] = num / sden;
    }
}

__global__ void depth_regression_kernel(const float* __restrict__ p, const float* __restrict__ dv, int is_map,
                                        float* __restrict__ out, int D, int64_t hw, int64_t total) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int64_t b = i / hw, pix = i % hw;
    float s = 0.0f;
    for (int k = 0; k < D; ++k) {
        const float d = is_map ? __ldg(dv + (b * D + k) * hw + pix) : __ldg(dv + b * D + k);
        s += __ldg(p + (b * D + k) * hw + pix) * d;
    }
    out[i] = s;
}

__global__ void conf_regression_kernel(const float* __restrict__ p, int n, float* __restrict__ out, int D, int64_t hw,
                                       int64_t total) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int64_t b = i / hw, pix = i % hw;
    const float* col = p + b * D * hw + pix;
    float e = 0.0f;
    for (int k = 0; k < D; ++k) e += __ldg(col + k * hw) * (float)k;
    int idx = (int)e;                       // .long() truncates toward zero
    idx = max(0, min(D - 1, idx));
    const int left = (n % 2 == 1) ? n / 2 : n / 2 - 1;
    float s = 0.0f;
    for (int t = 0; t < n; ++t) {
        const int k = idx - left + t;
        if (k >= 0 && k < D) s += __ldg(col + k * hw);
    }
    // the reference computes n * avg_pool(window): sum * (1/n) * n; reproduce the rounding
    out[i] = (float)n * (s / (float)n);
}

__global__ void init_range_kernel(const float* __restrict__ cur, int ND, int inverse, float* __restrict__ out, int D,
                                  int64_t hw, int64_t total) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int64_t bk = i / hw;
    const int b = (int)(bk / D), k = (int)(bk % D);
    const float d0 = __ldg(cur + (int64_t)b * ND), d1 = __ldg(cur + (int64_t)b * ND + ND - 1);
    if (inverse) {
        const float inv_near = 1.0f / d0, inv_far = 1.0f / d1;
        const float frac = (float)k / (float)(D - 1);
        out[i] = 1.0f / (inv_far + (inv_near - inv_far) * frac);
    } else {
        const float itv = (d1 - d0) / (float)(D - 1);
        out[i] = d0 + (float)k * itv;
    }
}

struct Lerp {
    int i0, i1;
    float l0, l1;
};

// ATen upsample (align_corners=True): src = dst * (in-1)/(out-1); i0 = (int)src; lambda1 = src - i0.
__device__ __forceinline__ Lerp lerp_axis(int o, int n_in, int n_out) {
    Lerp r;
    const float scale = n_out > 1 ? (float)(n_in - 1) / (float)(n_out - 1) : 0.0f;
    const float src = scale * (float)o;
    r.i0 = (int)src;
    r.i1 = r.i0 + ((r.i0 < n_in - 1) ? 1 : 0);
    r.l1 = src - (float)r.i0;
    r.l0 = 1.0f - r.l1;
    return r;
}

// One thread per output pixel; produces the whole column.  inverse = 1: module.py:642-653,
// inverse = 0: module.py:687-699.
__global__ void __launch_bounds__(256)
schedule_kernel(const float* __restrict__ depth, const float* __restrict__ hypo, int Dprev, float split_itv,
                const float* __restrict__ interval, int inverse, float* __restrict__ out, int D, int H, int W) {
    // 32 x 8 pixel blocks, grid.z = batch item: no 64-bit division to find (b, y, x)
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= W || y >= H) return;
    const int64_t b = blockIdx.z;
    const int h2 = H / 2, w2 = W / 2;
    const Lerp ly = lerp_axis(y, h2, H), lx = lerp_axis(x, w2, W);
    float lo[4], hi[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const int yy = (c & 2) ? ly.i1 : ly.i0, xx = (c & 1) ? lx.i1 : lx.i0;
        const int64_t q = (int64_t)yy * w2 + xx;
        const float dep = __ldg(depth + b * h2 * w2 + q);
        if (inverse) {
            const float itv = 1.0f / __ldg(hypo + (b * Dprev + 2) * h2 * w2 + q) - 1.0f / __ldg(hypo + (b * Dprev + 1) * h2 * w2 + q);
            const float inv = 1.0f / dep;
            hi[c] = inv + split_itv * itv;     // inverse_min_depth (near, k = D-1)
            lo[c] = inv - split_itv * itv;     // inverse_max_depth (far,  k = 0)
        } else {
            const float half = (float)D / 2.0f * __ldg(interval + b);
            lo[c] = fmaxf(dep - half, 0.01f);
            hi[c] = dep + half;
        }
    }
    const int64_t hw = (int64_t)H * W;
    float* o = out + b * D * hw + (int64_t)y * W + x;
    for (int k = 0; k < D; ++k) {
        float v[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if (inverse) {
                const float frac = (float)k / (float)(D - 1);
                v[c] = lo[c] + (hi[c] - lo[c]) * frac;
            } else {
                const float itv = (hi[c] - lo[c]) / (float)(D - 1);
                v[c] = lo[c] + (float)k * itv;
            }
        }
        const float r = ly.l0 * (lx.l0 * v[0] + lx.l1 * v[1]) + ly.l1 * (lx.l0 * v[2] + lx.l1 * v[3]);
        o[k * hw] = inverse ? 1.0f / r : r;
    }
}

__global__ void confidence_accumulate_kernel(const float* __restrict__ conf, int h, int w, float* __restrict__ acc,
                                             float* __restrict__ up, int H, int W, float scale) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= W || y >= H) return;
    const int64_t b = blockIdx.z, o = (b * H + y) * (int64_t)W + x;
    // F.interpolate(mode='nearest'): src = floor(dst * in/out) computed in fp32
    const int sy = min((int)floorf((float)y * ((float)h / (float)H)), h - 1);
    const int sx = min((int)floorf((float)x * ((float)w / (float)W)), w - 1);
    const float c = __ldg(conf + (b * h + sy) * (int64_t)w + sx);
    acc[o] += scale * c;
    if (up) up[o] = c;                          // the stage's own confidence, nearest-upsampled (:439-441)
}

static int check_map_args(const char* fn, const void* a, const void* b, const void* c, int B, int D, int H, int W) {
    MVS_REQUIRE(a && b && c, "%s: null pointer", fn);
    MVS_REQUIRE(B >= 1 && D >= 1 && H >= 1 && W >= 1, "%s: empty shape B=%d D=%d H=%d W=%d", fn, B, D, H, W);
    return MVS_OK;
}

}  // namespace mvs

extern "C" int mvs_prob_conv_cl(const float* x, const float* w_host, const float* bias_host, float* pre, int B, int D,
                                int H, int W, int Cin, int ksize, void* stream) {
    using namespace mvs;
    int rc = check_map_args("mvs_prob_conv_cl", x, w_host, pre, B, D, H, W);
    if (rc) return rc;
    if (Cin != 8) MVS_UNSUPPORTED("mvs_prob_conv_cl: only Cin = 8 is built (got %d)", Cin);
    MVS_REQUIRE(ksize == 1 || ksize == 3, "mvs_prob_conv_cl: ksize must be 1 or 3 (got %d)", ksize);
    ProbWeights P;
    memset(&P, 0, sizeof(P));
    memcpy(P.w, w_host, sizeof(float) * 8 * (ksize == 1 ? 1 : 27));
    P.bias = bias_host ? bias_host[0] : 0.0f;
    const int64_t total = (int64_t)B * D * H * W;
    if (ksize == 1)
        prob_conv1_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(x, pre, total, P);
    else {
        const int zchunks = cdiv(D, PZ);
        MVS_REQUIRE((int64_t)B * zchunks <= 65535 && cdiv(H, PT_Y) <= 65535, "mvs_prob_conv_cl: volume too large for one launch");
        dim3 grid(cdiv(W, PT_X), cdiv(H, PT_Y), B * zchunks);
        prob_conv3_tiled_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, pre, D, H, W, zchunks, P);
    }
    MVS_LAUNCH_OK("prob_conv_kernel");
    return MVS_OK;
}

extern "C" int mvs_regression_head(const float* pre, const float* depth_values, float tmp, int mode, float* prob_volume,
                                   float* depth, float* confidence, int B, int D, int H, int W, void* stream) {
    using namespace mvs;
    int rc = check_map_args("mvs_regression_head", pre, depth_values, depth, B, D, H, W);
    if (rc) return rc;
    MVS_REQUIRE(confidence, "mvs_regression_head: null confidence output");
    MVS_REQUIRE(mode == 0 || mode == 1, "mvs_regression_head: mode must be 0 (eval) or 1 (train), got %d", mode);
    const int64_t hw = (int64_t)H * W;
    MVS_REQUIRE(B <= 65535, "mvs_regression_head: batch %d too large for one launch", B);
    regression_head_kernel<<<dim3(cdiv(hw, 256), B), 256, 0, (cudaStream_t)stream>>>(pre, depth_values, tmp, mode, prob_volume,
                                                                                    depth, confidence, D, hw);
    MVS_LAUNCH_OK("regression_head_kernel");
    return MVS_OK;
}

extern "C" int mvs_depth_regression(const float* p, const float* depth_values, int depth_is_map, float* out, int B, int D,
                                    int H, int W, void* stream) {
    using namespace mvs;
    int rc = check_map_args("mvs_depth_regression", p, depth_values, out, B, D, H, W);
    if (rc) return rc;
    const int64_t hw = (int64_t)H * W, total = hw * B;
    depth_regression_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(p, depth_values, depth_is_map, out, D, hw, total);
    MVS_LAUNCH_OK("depth_regression_kernel");
    return MVS_OK;
}

extern "C" int mvs_conf_regression(const float* p, int n, float* out, int B, int D, int H, int W, void* stream) {
    using namespace mvs;
    int rc = check_map_args("mvs_conf_regression", p, p, out, B, D, H, W);
    if (rc) return rc;
    MVS_REQUIRE(n >= 1 && n <= D, "mvs_conf_regression: window n = %d out of range", n);
    const int64_t hw = (int64_t)H * W, total = hw * B;
    conf_regression_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(p, n, out, D, hw, total);
    MVS_LAUNCH_OK("conf_regression_kernel");
    return MVS_OK;
}

static int init_range_common(const char* fn, const float* cur_depth, int ND, int inverse, float* out, int B, int D, int H,
                             int W, void* stream) {
    using namespace mvs;
    int rc = check_map_args(fn, cur_depth, cur_depth, out, B, D, H, W);
    if (rc) return rc;
    MVS_REQUIRE(ND >= 2 && D >= 2, "%s: need at least 2 depth values and 2 hypotheses", fn);
    const int64_t hw = (int64_t)H * W, total = hw * B * D;
    init_range_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(cur_depth, ND, inverse, out, D, hw, total);
    MVS_LAUNCH_OK("init_range_kernel");
    return MVS_OK;
}

extern "C" int mvs_init_inverse_range(const float* cur_depth, int ND, float* out, int B, int D, int H, int W, void* stream) {
    return init_range_common("mvs_init_inverse_range", cur_depth, ND, 1, out, B, D, H, W, stream);
}

extern "C" int mvs_init_range(const float* cur_depth, int ND, float* out, int B, int D, int H, int W, void* stream) {
    return init_range_common("mvs_init_range", cur_depth, ND, 0, out, B, D, H, W, stream);
}

extern "C" int mvs_schedule_inverse_range(const float* depth, const float* depth_hypo, int Dprev, float split_itv,
                                          float* out, int B, int D, int H, int W, void* stream) {
    using namespace mvs;
    int rc = check_map_args("mvs_schedule_inverse_range", depth, depth_hypo, out, B, D, H, W);
    if (rc) return rc;
    MVS_REQUIRE(Dprev >= 3, "mvs_schedule_inverse_range: previous stage needs >= 3 hypotheses (got %d)", Dprev);
    MVS_REQUIRE(D >= 2 && H >= 2 && W >= 2, "mvs_schedule_inverse_range: D, H, W must be >= 2");
    MVS_REQUIRE(B <= 65535 && cdiv(H, 8) <= 65535, "mvs_schedule_inverse_range: map too large for one launch");
    schedule_kernel<<<dim3(cdiv(W, 32), cdiv(H, 8), B), 256, 0, (cudaStream_t)stream>>>(depth, depth_hypo, Dprev, split_itv, nullptr, 1,
                                                                                       out, D, H, W);
    MVS_LAUNCH_OK("schedule_kernel");
    return MVS_OK;
}

extern "C" int mvs_schedule_range(const float* depth, const float* interval, float* out, int B, int D, int H, int W,
                                  void* stream) {
    using namespace mvs;
    int rc = check_map_args("mvs_schedule_range", depth, interval, out, B, D, H, W);
    if (rc) return rc;
    MVS_REQUIRE(D >= 2 && H >= 2 && W >= 2, "mvs_schedule_range: D, H, W must be >= 2");
    MVS_REQUIRE(B <= 65535 && cdiv(H, 8) <= 65535, "mvs_schedule_range: map too large for one launch");
    schedule_kernel<<<dim3(cdiv(W, 32), cdiv(H, 8), B), 256, 0, (cudaStream_t)stream>>>(depth, nullptr, 0, 0.0f, interval, 0, out, D, H, W);
    MVS_LAUNCH_OK("schedule_kernel");
    return MVS_OK;
}

extern "C" int mvs_confidence_accumulate(const float* conf, int h, int w, float* acc, int B, int H, int W, float scale,
                                         void* stream) {
    using namespace mvs;
    MVS_REQUIRE(conf && acc, "mvs_confidence_accumulate: null pointer");
    MVS_REQUIRE(B >= 1 && h >= 1 && w >= 1 && H >= 1 && W >= 1, "mvs_confidence_accumulate: empty shape");
    MVS_REQUIRE(B <= 65535 && cdiv(H, 8) <= 65535, "mvs_confidence_accumulate: map too large for one launch");
    confidence_accumulate_kernel<<<dim3(cdiv(W, 32), cdiv(H, 8), B), 256, 0, (cudaStream_t)stream>>>(conf, h, w, acc, nullptr, H, W, scale);
    MVS_LAUNCH_OK("confidence_accumulate_kernel");
    return MVS_OK;
}

extern "C" int mvs_confidence_upsample_accumulate(const float* conf, int h, int w, float* up, float* acc, int B, int H, int W,
                                                  float scale, void* stream) {
    using namespace mvs;
    MVS_REQUIRE(conf && acc && up, "mvs_confidence_upsample_accumulate: null pointer");
    MVS_REQUIRE(B >= 1 && h >= 1 && w >= 1 && H >= 1 && W >= 1, "mvs_confidence_upsample_accumulate: empty shape");
    MVS_REQUIRE(B <= 65535 && cdiv(H, 8) <= 65535, "mvs_confidence_upsample_accumulate: map too large for one launch");
    confidence_accumulate_kernel<<<dim3(cdiv(W, 32), cdiv(H, 8), B), 256, 0, (cudaStream_t)stream>>>(conf, h, w, acc, up, H, W, scale);
    MVS_LAUNCH_OK("confidence_accumulate_kernel");
    return MVS_OK;
}
