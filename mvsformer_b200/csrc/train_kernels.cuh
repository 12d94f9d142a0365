// train_kernels.cuh — per-thread bodies of the training-path kernels (train-mode forward with
// batch-statistics BatchNorm, and the backward of the whole StageNet; SURVEY.md §8b "Autograd",
// BASELINE cfg 5).
//
// Every kernel of this file is "flat": thread `tid` of `nthreads` does its work with no
// shared memory, no warp collectives and no barriers; cross-thread reductions go through
// atomics: fp32 for feature and weight gradients (as ATen's grid_sampler / cuDNN's backward
// do), fp64 for the BatchNorm statistics and the BatchNorm gradient sums.
// That makes each body an ordinary function of (arguments, tid), which the test-suite's CPU
// emulation harness (tests/emu/emu.cpp) compiles from THIS source and runs thread by thread to
// check index arithmetic and formulas against torch autograd without a GPU.  train.cu wraps the
// bodies in __global__ kernels; nothing here is reachable from the product on a CPU.
//
// The includer provides: __device__, __forceinline__, __restrict__, __ldg, the *_rn intrinsics,
// MVS_ATOMIC_ADD_F(ptr, v), MVS_ATOMIC_ADD_D(ptr, v), <stdint.h>, <math.h>, and geometry.cuh.
//
// Reference lines (in /root/reference): models/mvsformer_model.py:61-105 (cost volume),
// models/warping.py:69-109 (warp; the grid is built under no_grad, :79, so only the sampled
// features receive gradients), models/module.py:83-197 (conv + BN + ReLU blocks),
// torch.nn.BatchNorm{2,3}d training semantics (biased batch variance for normalisation,
// unbiased for running_var, momentum 0.1).
#pragma once

namespace mvs {
namespace train {

// ------------------------------------------------------------------------------------------
// Group-wise correlation of ONE reference view with N source views, materialised per view
// (training only; the inference path never stores it).       mvsformer_model.py:70-79
//   corr[b][v][k][y][x][g] = (1/cpg) sum_c' ref[b][g*cpg+c'][y][x] * warped_v[b][g*cpg+c'][k][y][x]
// one thread per (b, v, k, y, x)
// ------------------------------------------------------------------------------------------
struct CorrDims {
    int B, V, C, G, D, H, W;
    int64_t batch_stride, view_stride;   // of the [B,V,C,H,W] feature tensor, in elements
};

__device__ __forceinline__ void decode_bvkyx(int64_t tid, const CorrDims& d, int* b, int* v, int* k, int* y, int* x) {
    *x = (int)(tid % d.W); tid /= d.W;
    *y = (int)(tid % d.H); tid /= d.H;
    *k = (int)(tid % d.D); tid /= d.D;
    const int n = d.V - 1;
    *v = (int)(tid % n);
    *b = (int)(tid / n);
}

__device__ __forceinline__ void group_corr_fwd_thread(const float* __restrict__ features, const float* __restrict__ relproj,
                                                      const float* __restrict__ depth, float* __restrict__ corr,
                                                      CorrDims d, int64_t tid) {
    const int n = d.V - 1;
    const int64_t total = (int64_t)d.B * n * d.D * d.H * d.W;
    if (tid >= total) return;
    int b, v, k, y, x;
    decode_bvkyx(tid, d, &b, &v, &k, &y, &x);
    const int64_t hw = (int64_t)d.H * d.W;
    const RelProj m = load_relproj(relproj + ((int64_t)b * n + v) * 12);
    const PixelRay ray = pixel_ray(m, (float)x, (float)y);
    const float dep = __ldg(depth + (((int64_t)b * d.D + k) * d.H + y) * d.W + x);
    const Taps t = make_taps(m, ray, dep, d.H, d.W, (float)(d.W - 1) / 2.0f, (float)(d.H - 1) / 2.0f);
    const float* ref = features + (int64_t)b * d.batch_stride + (int64_t)y * d.W + x;
    const float* src = features + (int64_t)b * d.batch_stride + (int64_t)(v + 1) * d.view_stride;
    const int cpg = d.C / d.G;
    const float inv_cpg = 1.0f / (float)cpg;
    float* out = corr + tid * d.G;           // [b][v][k][y][x][g] has the same linear order as tid
    for (int g = 0; g < d.G; ++g) {
        float acc = 0.0f;
        for (int c = g * cpg; c < (g + 1) * cpg; ++c)
            acc = fmaf(__ldg(ref + (int64_t)c * hw), sample4(src + (int64_t)c * hw, t), acc);
        out[g] = acc * inv_cpg;
    }
}

// Backward of the above w.r.t. the features (reference AND source views): one thread per
// (b, v, k, y, x); gfeat [B,V,C,H,W] dense, zero-initialised by the caller, accumulated with
// fp32 atomics.
__device__ __forceinline__ void group_corr_bwd_thread(const float* __restrict__ features, const float* __restrict__ relproj,
                                                      const float* __restrict__ depth, const float* __restrict__ gcorr,
                                                      float* __restrict__ gfeat, CorrDims d, int64_t tid) {
    const int n = d.V - 1;
    const int64_t total = (int64_t)d.B * n * d.D * d.H * d.W;
    if (tid >= total) return;
    int b, v, k, y, x;
    decode_bvkyx(tid, d, &b, &v, &k, &y, &x);
    const int64_t hw = (int64_t)d.H * d.W;
    const RelProj m = load_relproj(relproj + ((int64_t)b * n + v) * 12);
    const PixelRay ray = pixel_ray(m, (float)x, (float)y);
    const float dep = __ldg(depth + (((int64_t)b * d.D + k) * d.H + y) * d.W + x);
    const Taps t = make_taps(m, ray, dep, d.H, d.W, (float)(d.W - 1) / 2.0f, (float)(d.H - 1) / 2.0f);
    const float* ref = features + (int64_t)b * d.batch_stride + (int64_t)y * d.W + x;
    const float* src = features + (int64_t)b * d.batch_stride + (int64_t)(v + 1) * d.view_stride;
    const int64_t gb = (int64_t)b * d.V * d.C * hw;                 // gfeat is dense
    float* gref = gfeat + gb + (int64_t)y * d.W + x;
    float* gsrc = gfeat + gb + (int64_t)(v + 1) * d.C * hw;
    const int cpg = d.C / d.G;
    const float inv_cpg = 1.0f / (float)cpg;
    const float* gin = gcorr + tid * d.G;
    for (int g = 0; g < d.G; ++g) {
        const float gv = __ldg(gin + g) * inv_cpg;
        if (gv == 0.0f) continue;
        for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
            const float warped = sample4(src + (int64_t)c * hw, t);
            MVS_ATOMIC_ADD_F(gref + (int64_t)c * hw, gv * warped);
            const float gw = gv * __ldg(ref + (int64_t)c * hw);
            float* plane = gsrc + (int64_t)c * hw;
            if (t.w00 != 0.0f) MVS_ATOMIC_ADD_F(plane + t.o00, gw * t.w00);
            if (t.w01 != 0.0f) MVS_ATOMIC_ADD_F(plane + t.o01, gw * t.w01);
            if (t.w10 != 0.0f) MVS_ATOMIC_ADD_F(plane + t.o10, gw * t.w10);
            if (t.w11 != 0.0f) MVS_ATOMIC_ADD_F(plane + t.o11, gw * t.w11);
        }
    }
}

// Entropy of softmax_k(sum_g corr) per (b, v, y, x).          mvsformer_model.py:87-90
// one thread per (b, v, y, x); three passes over the depth column (max, normaliser, entropy)
__device__ __forceinline__ void corr_entropy_thread(const float* __restrict__ corr, float* __restrict__ entropy,
                                                    int BN, int G, int D, int H, int W, int64_t tid) {
    const int64_t hw = (int64_t)H * W;
    if (tid >= (int64_t)BN * hw) return;
    const int64_t pix = tid % hw, bv = tid / hw;
    const float* col = corr + (bv * D * hw + pix) * G;            // + k*hw*G
    float mx = -INFINITY;
    for (int k = 0; k < D; ++k) {
        const float* p = col + (int64_t)k * hw * G;
        float s = 0.0f;
        for (int g = 0; g < G; ++g) s += __ldg(p + g);
        mx = fmaxf(mx, s);
    }
    float z = 0.0f;
    for (int k = 0; k < D; ++k) {
        const float* p = col + (int64_t)k * hw * G;
        float s = 0.0f;
        for (int g = 0; g < G; ++g) s += __ldg(p + g);
        z += expf(s - mx);
    }
    float e = 0.0f;
    for (int k = 0; k < D; ++k) {
        const float* p = col + (int64_t)k * hw * G;
        float s = 0.0f;
        for (int g = 0; g < G; ++g) s += __ldg(p + g);
        const float pr = expf(s - mx) / z;
        e -= pr * logf(pr + 1e-7f);
    }
    entropy[tid] = e;
}

// Visibility-weighted aggregation over the source views.       mvsformer_model.py:101-105
//   volume[b][k][y][x][g] = (sum_v corr_v * w_v) / (sum_v w_v + 1e-6), views in order
// one thread per output element
__device__ __forceinline__ void aggregate_fwd_thread(const float* __restrict__ corr, const float* __restrict__ weight,
                                                     float* __restrict__ volume, int B, int N, int G, int D, int H, int W,
                                                     int64_t tid) {
    const int64_t hw = (int64_t)H * W;
    const int64_t per_b = (int64_t)D * hw * G;
    if (tid >= (int64_t)B * per_b) return;
    const int b = (int)(tid / per_b);
    const int64_t r = tid % per_b;                                   // (k, y, x, g)
    const int64_t pix = (r / G) % hw;
    float vol = 0.0f, vis = 0.0f;
    for (int v = 0; v < N; ++v) {
        const float w = __ldg(weight + ((int64_t)b * N + v) * hw + pix);
        vol = __fadd_rn(vol, __fmul_rn(__ldg(corr + ((int64_t)b * N + v) * per_b + r), w));
        vis = __fadd_rn(vis, w);
    }
    volume[tid] = __fdiv_rn(vol, __fadd_rn(vis, 1e-6f));
}

// Backward of the aggregation: one thread per (b, y, x).
//   gcorr_v = gvol * w_v / (S + eps);   gweight_v = sum_{k,g} gvol * (corr_v - vol) / (S + eps)
#define MVS_TRAIN_MAX_VIEWS 16
__device__ __forceinline__ void aggregate_bwd_thread(const float* __restrict__ gvol, const float* __restrict__ corr,
                                                     const float* __restrict__ weight, float* __restrict__ gcorr,
                                                     float* __restrict__ gweight, int B, int N, int G, int D, int H, int W,
                                                     int64_t tid) {
    const int64_t hw = (int64_t)H * W;
    if (tid >= (int64_t)B * hw) return;
    const int b = (int)(tid / hw);
    const int64_t pix = tid % hw;
    const int64_t per_b = (int64_t)D * hw * G;
    float w[MVS_TRAIN_MAX_VIEWS], gw[MVS_TRAIN_MAX_VIEWS];
    float vis = 0.0f;
    for (int v = 0; v < N; ++v) {
        w[v] = __ldg(weight + ((int64_t)b * N + v) * hw + pix);
        vis += w[v];
        gw[v] = 0.0f;
    }
    const float inv = 1.0f / (vis + 1e-6f);
    for (int k = 0; k < D; ++k) {
        for (int g = 0; g < G; ++g) {
            const int64_t r = ((int64_t)k * hw + pix) * G + g;
            const float gv = __ldg(gvol + (int64_t)b * per_b + r) * inv;
            float a = 0.0f;
            for (int v = 0; v < N; ++v) a = fmaf(__ldg(corr + ((int64_t)b * N + v) * per_b + r), w[v], a);
            const float vol = a * inv;
            for (int v = 0; v < N; ++v) {
                const int64_t o = ((int64_t)b * N + v) * per_b + r;
                gcorr[o] = gv * w[v];
                gw[v] = fmaf(gv, __ldg(corr + o) - vol, gw[v]);
            }
        }
    }
    for (int v = 0; v < N; ++v) gweight[((int64_t)b * N + v) * hw + pix] = gw[v];
}

// ------------------------------------------------------------------------------------------
// BatchNorm (training) over channels-last data x[M][C]:  nthreads % C == 0, so a thread always
// sees channel tid % C and consecutive threads read consecutive floats.
// ------------------------------------------------------------------------------------------
// The fp64 atomics of a reduction are spread over MVS_BN_REPLICAS copies of the 2C accumulators
// (replica = (tid / C) % R, so the 32 lanes of a warp hit 32 different addresses); bn_collapse
// folds the copies into replica 0, which is what finalize / apply / the host read.  Measured on
// B200 (profiles/r01_train_step_entry_points.json): the un-replicated version spent 10 of 41 ms of
// a training step in same-address fp64 atomics.
#define MVS_BN_REPLICAS 32
__device__ __forceinline__ void bn_stats_thread(const float* __restrict__ x, double* __restrict__ sums, int64_t M, int C,
                                                int64_t tid, int64_t nthreads) {
    const int64_t total = M * C;
    double s = 0.0, q = 0.0;
    for (int64_t e = tid; e < total; e += nthreads) {
        const double v = (double)__ldg(x + e);
        s += v;
        q += v * v;
    }
    if (tid < total) {
        const int c = (int)(tid % C);
        double* rep = sums + ((tid / C) % MVS_BN_REPLICAS) * 2 * C;
        MVS_ATOMIC_ADD_D(rep + c, s);
        MVS_ATOMIC_ADD_D(rep + C + c, q);
    }
}

// sums[0][i] += sum_{r >= 1} sums[r][i], one thread per i < 2C
__device__ __forceinline__ void bn_collapse_thread(double* __restrict__ sums, int C, int64_t tid) {
    if (tid >= 2 * C) return;
    double t = sums[tid];
    for (int r = 1; r < MVS_BN_REPLICAS; ++r) t += sums[(int64_t)r * 2 * C + tid];
    sums[tid] = t;
}

// sums[2C] (sum, sum of squares over `count` samples) -> mean_invstd[2C]; updates the running
// statistics in place when given (momentum m: r = (1-m) r + m * batch, unbiased batch variance).
// `replicas` = how many copies of the accumulators to fold (MVS_BN_REPLICAS straight after mvs_bn_stats,
// 1 after an explicit mvs_bn_collapse, e.g. around the SyncBatchNorm all-reduce).
__device__ __forceinline__ void bn_finalize_thread(const double* __restrict__ sums, int replicas, double count, float eps,
                                                   float momentum, float* __restrict__ mean_invstd,
                                                   float* __restrict__ running_mean, float* __restrict__ running_var, int C,
                                                   int64_t tid) {
    if (tid >= C) return;
    const int c = (int)tid;
    double s = 0.0, q = 0.0;
    for (int r = 0; r < replicas; ++r) { s += sums[(int64_t)r * 2 * C + c]; q += sums[(int64_t)r * 2 * C + C + c]; }
    const double mean = s / count;
    double var = q / count - mean * mean;
    if (var < 0.0) var = 0.0;
    mean_invstd[c] = (float)mean;
    mean_invstd[C + c] = (float)(1.0 / sqrt(var + (double)eps));
    if (running_mean) running_mean[c] = (float)((1.0 - (double)momentum) * (double)running_mean[c] + (double)momentum * mean);
    if (running_var) {
        const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
        running_var[c] = (float)((1.0 - (double)momentum) * (double)running_var[c] + (double)momentum * unbiased);
    }
}

// y = act((x - mean) * invstd * gamma + beta) (+ skip), one thread per element
__device__ __forceinline__ void bn_act_fwd_thread(const float* __restrict__ x, const float* __restrict__ mean_invstd,
                                                  const float* __restrict__ gamma, const float* __restrict__ beta,
                                                  const float* __restrict__ skip, float* __restrict__ y, int64_t M, int C,
                                                  int relu, int64_t tid) {
    if (tid >= M * C) return;
    const int c = (int)(tid % C);
    const float xh = (__ldg(x + tid) - __ldg(mean_invstd + c)) * __ldg(mean_invstd + C + c);
    float o = fmaf(xh, __ldg(gamma + c), __ldg(beta + c));
    if (relu) o = fmaxf(o, 0.0f);
    if (skip) o += __ldg(skip + tid);
    y[tid] = o;
}

// sums[0..C) += sum g * xhat (d gamma), sums[C..2C) += sum g (d beta), g = gy masked by the ReLU
__device__ __forceinline__ void bn_act_bwd_reduce_thread(const float* __restrict__ gy, const float* __restrict__ x,
                                                         const float* __restrict__ mean_invstd,
                                                         const float* __restrict__ gamma, const float* __restrict__ beta,
                                                         double* __restrict__ sums, int64_t M, int C, int relu, int64_t tid,
                                                         int64_t nthreads) {
    const int64_t total = M * C;
    if (tid >= total) return;
    const int c = (int)(tid % C);
    const float mean = __ldg(mean_invstd + c), invstd = __ldg(mean_invstd + C + c);
    const float ga = __ldg(gamma + c), be = __ldg(beta + c);
    double dg = 0.0, db = 0.0;
    for (int64_t e = tid; e < total; e += nthreads) {
        const float xh = (__ldg(x + e) - mean) * invstd;
        float g = __ldg(gy + e);
        if (relu && !(fmaf(xh, ga, be) > 0.0f)) g = 0.0f;
        dg += (double)g * (double)xh;
        db += (double)g;
    }
    double* rep = sums + ((tid / C) % MVS_BN_REPLICAS) * 2 * C;
    MVS_ATOMIC_ADD_D(rep + c, dg);
    MVS_ATOMIC_ADD_D(rep + C + c, db);
}

// gx = gamma * invstd * (g - dbeta/count - xhat * dgamma/count), one thread per element
__device__ __forceinline__ void bn_act_bwd_apply_thread(const float* __restrict__ gy, const float* __restrict__ x,
                                                        const float* __restrict__ mean_invstd,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        const double* __restrict__ sums, double count,
                                                        float* __restrict__ gx, int64_t M, int C, int relu, int64_t tid) {
    if (tid >= M * C) return;
    const int c = (int)(tid % C);
    const float invstd = __ldg(mean_invstd + C + c);
    const float ga = __ldg(gamma + c);
    const float xh = (__ldg(x + tid) - __ldg(mean_invstd + c)) * invstd;
    float g = __ldg(gy + tid);
    if (relu && !(fmaf(xh, ga, __ldg(beta + c)) > 0.0f)) g = 0.0f;
    const float mdg = (float)(sums[c] / count), mdb = (float)(sums[C + c] / count);
    gx[tid] = ga * invstd * (g - mdb - xh * mdg);
}

// ------------------------------------------------------------------------------------------
// Weight gradient of a (kd,k,k) convolution / transposed convolution, padding k/2 per axis.
// `small` is the tensor on the strided (coarse) grid, `big` the one on the fine grid:
//   big position = small position * stride - pad + tap
// conv:    small = grad of the output (Cs = Cout), big = the input (Cb = Cin)   -> small_is_cout = 1
// deconv:  small = the input (Cs = Cin),  big = grad of the output (Cb = Cout)  -> small_is_cout = 0
// dw is the packed layout [kd][k][k][Cin][Cout] (zero-initialised by the caller, fp32 atomics).
// A thread owns one row (b, z, y) of the small grid x one tap x a TS x TB channel tile.
// ------------------------------------------------------------------------------------------
struct WgradDims {
    int B, Ds, Hs, Ws, Db, Hb, Wb, Cs, Cb;
    int kd, khw, sd, shw, small_is_cout;
    int rpt;   // rows of the small grid per thread (consecutive rows; fewer, fatter threads = fewer atomics)
    int xseg;  // x positions of a row per thread (rows are split into ceil(Ws / xseg) segments when there are
               // too few (row, task) pairs to fill the GPU); xseg >= Ws = whole rows
};

// T consecutive floats; 128-bit loads when T is a multiple of 4 (the caller guarantees 16-byte alignment:
// channel counts that are multiples of 4 and 16-byte aligned tensors)
template <int T>
__device__ __forceinline__ void load_vec(const float* __restrict__ p, float* out) {
    if (T % 4 == 0) {
#pragma unroll
        for (int q = 0; q < T / 4; ++q) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(p) + q);
            out[q * 4 + 0] = v.x; out[q * 4 + 1] = v.y; out[q * 4 + 2] = v.z; out[q * 4 + 3] = v.w;
        }
    } else {
#pragma unroll
        for (int i = 0; i < T; ++i) out[i] = __ldg(p + i);
    }
}

template <int TS, int TB>
__device__ __forceinline__ void conv_wgrad_thread(const float* __restrict__ small, const float* __restrict__ big,
                                                  float* __restrict__ dw, WgradDims d, int64_t tid) {
    const int nts = d.Cs / TS, ntb = d.Cb / TB;
    const int ntaps = d.kd * d.khw * d.khw;
    const int64_t ntasks = (int64_t)ntaps * nts * ntb;
    const int64_t nrows = (int64_t)d.B * d.Ds * d.Hs;
    const int64_t nchunks = (nrows + d.rpt - 1) / d.rpt;
    const int nseg = (d.Ws + d.xseg - 1) / d.xseg;
    if (tid >= ntasks * nchunks * nseg) return;
    int64_t task = tid % ntasks;
    const int64_t unit = tid / ntasks;
    const int seg = (int)(unit % nseg);
    const int64_t row0 = (unit / nseg) * d.rpt;
    const int tb = (int)(task % ntb); task /= ntb;
    const int ts = (int)(task % nts); task /= nts;
    const int tap = (int)task;
    const int kx = tap % d.khw, ky = (tap / d.khw) % d.khw, kz = tap / (d.khw * d.khw);
    // x range of this thread with the fine-grid position bx = x*shw - pad + kx inside [0, Wb): no
    // bounds test in the inner loop, so it unrolls and keeps several independent loads in flight
    // (measured on B200: the branchy one-load-at-a-time loop was latency bound, 13 of 41 ms per step)
    const int pad = d.khw / 2;
    int xa = (pad - kx > 0) ? 1 : 0;                              // ceil((pad - kx) / shw) for pad - kx in {-1, 0, 1}
    const int last = d.Wb - 1 + pad - kx;                         // largest admissible x*shw
    int xb = last < 0 ? 0 : last / d.shw + 1;
    if (xb > d.Ws) xb = d.Ws;
    if (xa < seg * d.xseg) xa = seg * d.xseg;
    if (xb > (seg + 1) * d.xseg) xb = (seg + 1) * d.xseg;
    if (xa >= xb) return;
    float acc[TS][TB];
#pragma unroll
    for (int i = 0; i < TS; ++i)
#pragma unroll
        for (int j = 0; j < TB; ++j) acc[i][j] = 0.0f;
    bool any = false;
    for (int64_t row = row0; row < row0 + d.rpt && row < nrows; ++row) {
        const int y = (int)(row % d.Hs);
        const int z = (int)((row / d.Hs) % d.Ds);
        const int b = (int)(row / ((int64_t)d.Hs * d.Ds));
        const int bz = z * d.sd - d.kd / 2 + kz;
        const int by = y * d.shw - pad + ky;
        if (bz < 0 || bz >= d.Db || by < 0 || by >= d.Hb) continue;
        any = true;
        const float* ps = small + ((((int64_t)b * d.Ds + z) * d.Hs + y) * d.Ws) * d.Cs + ts * TS;
        const float* pb = big + ((((int64_t)b * d.Db + bz) * d.Hb + by) * d.Wb + (kx - pad)) * d.Cb + tb * TB;
        const int64_t bstep = (int64_t)d.shw * d.Cb;
#pragma unroll 2
        for (int x = xa; x < xb; ++x) {
            float a[TS], c[TB];
            load_vec<TS>(ps + (int64_t)x * d.Cs, a);
            load_vec<TB>(pb + x * bstep, c);
#if defined(__CUDA_ARCH__)
            if (TB % 4 == 0) {                                   // packed FP32 (FFMA2): each half is the scalar fmaf below
#pragma unroll
                for (int i = 0; i < TS; ++i)
#pragma unroll
                    for (int j = 0; j < TB; j += 4) fma4_bcast(&acc[i][j], a[i], make_float4(c[j], c[j + 1], c[j + 2], c[j + 3]));
            } else
#endif
            {
#pragma unroll
                for (int i = 0; i < TS; ++i)
#pragma unroll
                    for (int j = 0; j < TB; ++j) acc[i][j] = fmaf(a[i], c[j], acc[i][j]);
            }
        }
    }
    if (!any) return;
    const int cin = d.small_is_cout ? d.Cb : d.Cs, cout = d.small_is_cout ? d.Cs : d.Cb;
    float* base = dw + (int64_t)tap * cin * cout;
#pragma unroll
    for (int i = 0; i < TS; ++i)
#pragma unroll
        for (int j = 0; j < TB; ++j) {
            const int cs = ts * TS + i, cb = tb * TB + j;
            const int ci = d.small_is_cout ? cb : cs, co = d.small_is_cout ? cs : cb;
            MVS_ATOMIC_ADD_F(base + (int64_t)ci * cout + co, acc[i][j]);
        }
}

// ------------------------------------------------------------------------------------------
// "Thin" convolution with device-resident weights for the few-channel layers (vis net 1->16,
// its 1x1 8->1 head, the 8->1 `prob` convs, and the data gradients of those): stride 1,
// padding k/2, kernel (kd,k,k); w [kd*k*k][Cin][Cout]; act 0 none, 1 ReLU, 2 sigmoid.
// one thread per (output voxel, output channel)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void thin_conv_thread(const float* __restrict__ x, const float* __restrict__ w,
                                                 const float* __restrict__ bias, float* __restrict__ y, int B, int D, int H,
                                                 int W, int Cin, int Cout, int kd, int khw, int act, int64_t tid) {
    const int64_t total = (int64_t)B * D * H * W * Cout;
    if (tid >= total) return;
    int64_t r = tid;
    const int co = (int)(r % Cout); r /= Cout;
    const int ox = (int)(r % W); r /= W;
    const int oy = (int)(r % H); r /= H;
    const int oz = (int)(r % D);
    const int b = (int)(r / D);
    float acc = bias ? __ldg(bias + co) : 0.0f;
    for (int kz = 0; kz < kd; ++kz) {
        const int iz = oz - kd / 2 + kz;
        if (iz < 0 || iz >= D) continue;
        for (int ky = 0; ky < khw; ++ky) {
            const int iy = oy - khw / 2 + ky;
            if (iy < 0 || iy >= H) continue;
            for (int kx = 0; kx < khw; ++kx) {
                const int ix = ox - khw / 2 + kx;
                if (ix < 0 || ix >= W) continue;
                const float* px = x + ((((int64_t)b * D + iz) * H + iy) * W + ix) * Cin;
                const float* pw = w + ((int64_t)((kz * khw + ky) * khw + kx) * Cin) * Cout + co;
                for (int ci = 0; ci < Cin; ++ci) acc = fmaf(__ldg(px + ci), __ldg(pw + (int64_t)ci * Cout), acc);
            }
        }
    }
    if (act == 1) acc = fmaxf(acc, 0.0f);
    else if (act == 2) acc = 1.0f / (1.0f + expf(-acc));
    y[tid] = acc;
}

// gx = gy * y * (1 - y)
__device__ __forceinline__ void sigmoid_bwd_thread(const float* __restrict__ gy, const float* __restrict__ y,
                                                   float* __restrict__ gx, int64_t n, int64_t tid) {
    if (tid >= n) return;
    const float s = __ldg(y + tid);
    gx[tid] = __ldg(gy + tid) * s * (1.0f - s);
}

// Backward of p = softmax_k(pre) along D of a [B,D,H,W] volume: gpre = p * (gp - sum_k gp * p)
// one thread per (b, y, x)
__device__ __forceinline__ void softmax_bwd_thread(const float* __restrict__ gp, const float* __restrict__ p,
                                                   float* __restrict__ gpre, int B, int D, int64_t hw, int64_t tid) {
    if (tid >= (int64_t)B * hw) return;
    const int64_t b = tid / hw, pix = tid % hw;
    const int64_t base = b * D * hw + pix;
    float dot = 0.0f;
    for (int k = 0; k < D; ++k) dot = fmaf(__ldg(gp + base + (int64_t)k * hw), __ldg(p + base + (int64_t)k * hw), dot);
    for (int k = 0; k < D; ++k) {
        const int64_t o = base + (int64_t)k * hw;
        gpre[o] = __ldg(p + o) * (__ldg(gp + o) - dot);
    }
}

// Backward of the materialising warp (mvs_homo_warp; models/warping.py:105 F.grid_sample) w.r.t.
// the source feature: gsrc[b][c][tap] += gwarped[b][c][k][y][x] * weight(tap).  gsrc [B,C,H,W]
// zero-initialised by the caller; one thread per (b, k, y, x).
__device__ __forceinline__ void homo_warp_bwd_thread(const float* __restrict__ gwarped, const float* __restrict__ relproj,
                                                     const float* __restrict__ depth, int depth_is_map,
                                                     float* __restrict__ gsrc, int B, int C, int D, int H, int W,
                                                     int64_t tid) {
    const int64_t hw = (int64_t)H * W;
    if (tid >= (int64_t)B * D * hw) return;
    const int x = (int)(tid % W), y = (int)((tid / W) % H);
    const int k = (int)((tid / hw) % D), b = (int)(tid / (hw * D));
    const RelProj m = load_relproj(relproj + (int64_t)b * 12);
    const PixelRay ray = pixel_ray(m, (float)x, (float)y);
    const float dep = depth_is_map ? __ldg(depth + ((int64_t)b * D + k) * hw + (int64_t)y * W + x)
                                   : __ldg(depth + (int64_t)b * D + k);
    const Taps t = make_taps(m, ray, dep, H, W, (float)((W - 1) / 2.0), (float)((H - 1) / 2.0));
    const float* gp = gwarped + (((int64_t)b * C) * D + k) * hw + (int64_t)y * W + x;
    float* plane = gsrc + (int64_t)b * C * hw;
    for (int c = 0; c < C; ++c, plane += hw) {
        const float g = __ldg(gp + (int64_t)c * D * hw);
        if (g == 0.0f) continue;
        if (t.w00 != 0.0f) MVS_ATOMIC_ADD_F(plane + t.o00, g * t.w00);
        if (t.w01 != 0.0f) MVS_ATOMIC_ADD_F(plane + t.o01, g * t.w01);
        if (t.w10 != 0.0f) MVS_ATOMIC_ADD_F(plane + t.o10, g * t.w10);
        if (t.w11 != 0.0f) MVS_ATOMIC_ADD_F(plane + t.o11, g * t.w11);
    }
}

// ------------------------------------------------------------------------------------------
// fusion_type 'epipole' / 'epipoleV2' (models/mvsformer_model.py:92-104): the view weight is a softmax over the depth
// hypotheses of the summed correlation, per pixel AND per hypothesis:
//   s_v[k] = sum_g corr_v[k,g] / T  (- 10000 * proj_mask_v[k] for V2);  w_v[k] = softmax_k(s_v)[k] / norm;
//   volume[k,g] = sum_v corr_v[k,g] w_v[k] / (sum_v w_v[k] + 1e-6)
// corr [B,N,D,H,W,G], mask [B,N,D,H,W] (1/0) or NULL, stats [B,N,H,W,2] = (max_k s, sum_k exp(s - max)).
// ------------------------------------------------------------------------------------------
// out-of-bounds / behind-the-camera flag of a sample (models/warping.py:99-103); one thread per (b, v, k, y, x)
__device__ __forceinline__ void proj_mask_thread(const float* __restrict__ relproj, const float* __restrict__ depth,
                                                 float* __restrict__ mask, int B, int N, int D, int H, int W, int64_t tid) {
    const int64_t hw = (int64_t)H * W;
    if (tid >= (int64_t)B * N * D * hw) return;
    const int x = (int)(tid % W), y = (int)((tid / W) % H);
    const int k = (int)((tid / hw) % D);
    const int64_t bv = tid / (hw * D);
    const int64_t b = bv / N;
    const RelProj m = load_relproj(relproj + bv * 12);
    const PixelRay ray = pixel_ray(m, (float)x, (float)y);
    const float dep = __ldg(depth + ((b * D + k) * H + y) * W + x);
    float gx, gy, qz;
    make_taps(m, ray, dep, H, W, (float)((W - 1) / 2.0), (float)((H - 1) / 2.0), &gx, &gy, &qz);
    mask[tid] = ((gx > 1.0f) || (gx < -1.0f) || (gy > 1.0f) || (gy < -1.0f) || (qz <= 0.0f)) ? 1.0f : 0.0f;
}

__device__ __forceinline__ float epipole_score(const float* __restrict__ p, int G, float inv_t, const float* __restrict__ mk) {
    float s = 0.0f;
    for (int g = 0; g < G; ++g) s += __ldg(p + g);
    s *= inv_t;
    if (mk) s += -10000.0f * __ldg(mk);
    return s;
}

// one thread per (b, v, y, x): softmax statistics of the depth column
__device__ __forceinline__ void epipole_stats_thread(const float* __restrict__ corr, const float* __restrict__ mask, float inv_t,
                                                     float* __restrict__ stats, int BN, int G, int D, int64_t hw, int64_t tid) {
    if (tid >= (int64_t)BN * hw) return;
    const int64_t pix = tid % hw, bv = tid / hw;
    float mx = -INFINITY;
    for (int k = 0; k < D; ++k)
        mx = fmaxf(mx, epipole_score(corr + ((bv * D + k) * hw + pix) * G, G, inv_t, mask ? mask + (bv * D + k) * hw + pix : nullptr));
    float z = 0.0f;
    for (int k = 0; k < D; ++k)
        z += expf(epipole_score(corr + ((bv * D + k) * hw + pix) * G, G, inv_t, mask ? mask + (bv * D + k) * hw + pix : nullptr) - mx);
    stats[tid * 2] = mx;
    stats[tid * 2 + 1] = z;
}

// one thread per (b, k, y, x): volume [B,D,H,W,G] and wsum [B,D,H,W]
__device__ __forceinline__ void epipole_aggregate_fwd_thread(const float* __restrict__ corr, const float* __restrict__ mask,
                                                             const float* __restrict__ stats, float inv_t, float inv_norm,
                                                             float* __restrict__ volume, float* __restrict__ wsum, int B, int N,
                                                             int G, int D, int64_t hw, int64_t tid) {
    if (tid >= (int64_t)B * D * hw) return;
    const int64_t pix = tid % hw;
    const int k = (int)((tid / hw) % D);
    const int64_t b = tid / (hw * D);
    float acc[8];
    for (int g = 0; g < G; ++g) acc[g] = 0.0f;
    float ws = 0.0f;
    for (int v = 0; v < N; ++v) {
        const int64_t bv = b * N + v;
        const float* p = corr + ((bv * D + k) * hw + pix) * G;
        const float s = epipole_score(p, G, inv_t, mask ? mask + (bv * D + k) * hw + pix : nullptr);
        const float w = expf(s - __ldg(stats + (bv * hw + pix) * 2)) / __ldg(stats + (bv * hw + pix) * 2 + 1) * inv_norm;
        for (int g = 0; g < G; ++g) acc[g] += __ldg(p + g) * w;
        ws += w;
    }
    const float inv = 1.0f / (ws + 1e-6f);
    for (int g = 0; g < G; ++g) volume[tid * G + g] = acc[g] * inv;
    wsum[tid] = ws;
}

// Backward, one thread per (b, v, y, x) (a whole depth column, because of the softmax):
//   direct:  gcorr[k,g]  = gvol[k,g] * w[k] / (S[k] + eps)
//   via w:   gw[k]       = sum_g gvol[k,g] (corr[k,g] - vol[k,g]) / (S[k] + eps)
//            gs[k]       = w[k] (gw[k] - sum_j gw[j] p[j]),  p = softmax = w * norm      (d w_k / d s_j = (p_k (delta - p_j)) / norm)
//            gcorr[k,g] += gs[k] / T;    gT -= sum_k gs[k] * (sum_g corr[k,g]) / T^2     (fp32 atomic on gtemp[0])
__device__ __forceinline__ void epipole_aggregate_bwd_thread(const float* __restrict__ gvol, const float* __restrict__ corr,
                                                             const float* __restrict__ mask, const float* __restrict__ stats,
                                                             const float* __restrict__ volume, const float* __restrict__ wsum,
                                                             float inv_t, float inv_norm, float* __restrict__ gcorr,
                                                             float* __restrict__ gtemp, int B, int N, int G, int D, int64_t hw,
                                                             int64_t tid) {
    if (tid >= (int64_t)B * N * hw) return;
    const int64_t pix = tid % hw, bv = tid / hw;
    const int64_t b = bv / N;
    const float mx = __ldg(stats + tid * 2), z = __ldg(stats + tid * 2 + 1);
    float dot = 0.0f;                                              // sum_j gw[j] p[j]
    for (int k = 0; k < D; ++k) {
        const float* p = corr + ((bv * D + k) * hw + pix) * G;
        const int64_t o = (b * D + k) * hw + pix;
        const float prob = expf(epipole_score(p, G, inv_t, mask ? mask + (bv * D + k) * hw + pix : nullptr) - mx) / z;
        const float inv = 1.0f / (__ldg(wsum + o) + 1e-6f);
        float gw = 0.0f;
        for (int g = 0; g < G; ++g) gw += __ldg(gvol + o * G + g) * (__ldg(p + g) - __ldg(volume + o * G + g));
        dot += gw * inv * prob;
    }
    float gt = 0.0f;
    for (int k = 0; k < D; ++k) {
        const float* p = corr + ((bv * D + k) * hw + pix) * G;
        const int64_t o = (b * D + k) * hw + pix;
        const float prob = expf(epipole_score(p, G, inv_t, mask ? mask + (bv * D + k) * hw + pix : nullptr) - mx) / z;
        const float w = prob * inv_norm;
        const float inv = 1.0f / (__ldg(wsum + o) + 1e-6f);
        float gw = 0.0f, raw = 0.0f;
        for (int g = 0; g < G; ++g) {
            gw += __ldg(gvol + o * G + g) * (__ldg(p + g) - __ldg(volume + o * G + g));
            raw += __ldg(p + g);
        }
        const float gs = w * (gw * inv - dot);
        for (int g = 0; g < G; ++g) gcorr[((bv * D + k) * hw + pix) * G + g] = __ldg(gvol + o * G + g) * w * inv + gs * inv_t;
        gt -= gs * raw * inv_t * inv_t;
    }
    if (gtemp) MVS_ATOMIC_ADD_F(gtemp + (tid % 32), gt);          // 32 replicas, summed by the caller
}

// Backward of the warp through the SAMPLING GRID (diff_homo_warping_3D_with_mask, models/warping.py:112-152, where the
// grid is NOT under no_grad): gradients w.r.t. the depth hypotheses and the relative projection [R|t].
//   d warped / d ix = sum of the bilinear x-differences of the (zero-padded) taps, as ATen's grid_sampler backward;
//   ix = px, iy = py (the normalise / un-normalise round trip is the identity), p = q.xy / (q.z + 1e-6), q = R (x,y,1) d + t.
// gdepth [B,D,H,W] (depth_is_map) or [B,D] (atomics); grelproj [B, MVS_WARP_GRAD_REPLICAS, 12] (atomics spread over
// replicas, summed by the caller); both ZEROED by the caller.  One thread per (b, k, y, x).
#define MVS_WARP_GRAD_REPLICAS 32
__device__ __forceinline__ void homo_warp_bwd_grid_thread(const float* __restrict__ gwarped, const float* __restrict__ src,
                                                          const float* __restrict__ relproj, const float* __restrict__ depth,
                                                          int depth_is_map, float* __restrict__ gdepth,
                                                          float* __restrict__ grelproj, int B, int C, int D, int H, int W,
                                                          int64_t tid) {
    const int64_t hw = (int64_t)H * W;
    if (tid >= (int64_t)B * D * hw) return;
    const int x = (int)(tid % W), y = (int)((tid / W) % H);
    const int k = (int)((tid / hw) % D), b = (int)(tid / (hw * D));
    const RelProj m = load_relproj(relproj + (int64_t)b * 12);
    const PixelRay ray = pixel_ray(m, (float)x, (float)y);
    const float dep = depth_is_map ? __ldg(depth + ((int64_t)b * D + k) * hw + (int64_t)y * W + x)
                                   : __ldg(depth + (int64_t)b * D + k);
    const float qx = ray.x * dep + m.t0, qy = ray.y * dep + m.t1, qz = ray.z * dep + m.t2;
    const float den = qz + 1e-6f;
    const float ix = qx / den, iy = qy / den;
    const float x0 = floorf(ix), y0 = floorf(iy);
    const float wx1 = ix - x0, wx0 = 1.0f - wx1, wy1 = iy - y0, wy0 = 1.0f - wy1;
    const bool vx0 = x0 >= 0.0f && x0 <= (float)(W - 1), vx1 = x0 + 1.0f >= 0.0f && x0 + 1.0f <= (float)(W - 1);
    const bool vy0 = y0 >= 0.0f && y0 <= (float)(H - 1), vy1 = y0 + 1.0f >= 0.0f && y0 + 1.0f <= (float)(H - 1);
    if (!((vx0 || vx1) && (vy0 || vy1))) return;                      // every tap is padding (also NaN): zero gradient
    const int xi0 = vx0 ? (int)x0 : 0, xi1 = vx1 ? (int)x0 + 1 : 0, yi0 = vy0 ? (int)y0 : 0, yi1 = vy1 ? (int)y0 + 1 : 0;
    const float* sp = src + (int64_t)b * C * hw;
    const float* gp = gwarped + (((int64_t)b * C) * D + k) * hw + (int64_t)y * W + x;
    float gix = 0.0f, giy = 0.0f;
    for (int c = 0; c < C; ++c, sp += hw) {
        const float g = __ldg(gp + (int64_t)c * D * hw);
        const float s00 = (vx0 && vy0) ? __ldg(sp + (int64_t)yi0 * W + xi0) : 0.0f;
        const float s01 = (vx1 && vy0) ? __ldg(sp + (int64_t)yi0 * W + xi1) : 0.0f;
        const float s10 = (vx0 && vy1) ? __ldg(sp + (int64_t)yi1 * W + xi0) : 0.0f;
        const float s11 = (vx1 && vy1) ? __ldg(sp + (int64_t)yi1 * W + xi1) : 0.0f;
        gix += g * ((s01 - s00) * wy0 + (s11 - s10) * wy1);
        giy += g * ((s10 - s00) * wx0 + (s11 - s01) * wx1);
    }
    // p = q.xy / den
    const float gqx = gix / den, gqy = giy / den, gqz = -(gix * qx + giy * qy) / (den * den);
    const float gd = gqx * ray.x + gqy * ray.y + gqz * ray.z;
    if (depth_is_map) gdepth[((int64_t)b * D + k) * hw + (int64_t)y * W + x] = gd;
    else MVS_ATOMIC_ADD_F(gdepth + (int64_t)b * D + k, gd);
    float* gr = grelproj + ((int64_t)b * MVS_WARP_GRAD_REPLICAS + (tid % MVS_WARP_GRAD_REPLICAS)) * 12;
    const float px = (float)x * dep, py = (float)y * dep;              // d q / d R row = depth * (x, y, 1)
    MVS_ATOMIC_ADD_F(gr + 0, gqx * px); MVS_ATOMIC_ADD_F(gr + 1, gqx * py); MVS_ATOMIC_ADD_F(gr + 2, gqx * dep); MVS_ATOMIC_ADD_F(gr + 3, gqx);
    MVS_ATOMIC_ADD_F(gr + 4, gqy * px); MVS_ATOMIC_ADD_F(gr + 5, gqy * py); MVS_ATOMIC_ADD_F(gr + 6, gqy * dep); MVS_ATOMIC_ADD_F(gr + 7, gqy);
    MVS_ATOMIC_ADD_F(gr + 8, gqz * px); MVS_ATOMIC_ADD_F(gr + 9, gqz * py); MVS_ATOMIC_ADD_F(gr + 10, gqz * dep); MVS_ATOMIC_ADD_F(gr + 11, gqz);
}

// Backward of depth_regression (models/module.py:597-603): depth = sum_k p[k] * d[k]  ->  gp[k] = gdepth * d[k].
// d is [B,D,H,W] (depth_is_map) or [B,D]; one thread per (b, k, y, x)
__device__ __forceinline__ void depth_regression_bwd_thread(const float* __restrict__ gdepth, const float* __restrict__ dv,
                                                            int depth_is_map, float* __restrict__ gp, int B, int D,
                                                            int64_t hw, int64_t tid) {
    if (tid >= (int64_t)B * D * hw) return;
    const int64_t pix = tid % hw, bk = tid / hw;
    const int64_t b = bk / D;
    const float d = depth_is_map ? __ldg(dv + tid) : __ldg(dv + bk);
    gp[tid] = __ldg(gdepth + b * hw + pix) * d;
}

// depth_type = 'mixup_ce' head (models/mvsformer_model.py:126-136): adjacent-pair probabilities, first maximal pair,
// renormalised two-point depth.  prob, dv [B,D,H,W] -> depth, conf [B,H,W]; one thread per (b, y, x)
__device__ __forceinline__ void mixup_head_thread(const float* __restrict__ prob, const float* __restrict__ dv,
                                                  float* __restrict__ depth, float* __restrict__ conf, int B, int D,
                                                  int64_t hw, int64_t tid) {
    if (tid >= (int64_t)B * hw) return;
    const int64_t b = tid / hw, pix = tid % hw;
    const float* p = prob + b * D * hw + pix;
    const float* d = dv + b * D * hw + pix;
    float best = -INFINITY;
    int arg = 0;
    for (int k = 0; k + 1 < D; ++k) {
        const float m = __ldg(p + (int64_t)k * hw) + __ldg(p + (int64_t)(k + 1) * hw);
        if (m > best) { best = m; arg = k; }                      // torch.max: first maximal index
    }
    const float pl = __ldg(p + (int64_t)arg * hw), pr = __ldg(p + (int64_t)(arg + 1) * hw);
    const float s = pl + pr + 1e-7f;
    conf[tid] = best;
    depth[tid] = __ldg(d + (int64_t)arg * hw) * (pl / s) + __ldg(d + (int64_t)(arg + 1) * hw) * (pr / s);
}

// ------------------------------------------------------------------------------------------
// Functors: one per kernel, called as f(tid, nthreads) by launch_flat (train.cu: a __global__
// wrapper; tests/emu/emu.cpp: a loop over tid).
// ------------------------------------------------------------------------------------------
struct GroupCorrFwd {
    const float *features, *relproj, *depth; float* corr; CorrDims d;
    __device__ __forceinline__ void operator()(int64_t tid, int64_t) const { group_corr_fwd_thread(features, relproj, depth, corr, d, tid); }
};
struct GroupCorrBwd {
    const float *features, *relproj, *depth, *gcorr; float* gfeat; CorrDims d;
    __device__ __forceinline__ void operator()(int64_t tid, int64_t) const { group_corr_bwd_thread(features, relproj, depth, gcorr, gfeat, d, tid); }
};
struct CorrEntropy {
    const float* corr; float* entropy; int BN, G, D, H, W;
    __device__ __forceinline__ void operator()(int64_t tid, int64_t) const { corr_entropy_thread(corr, entropy, BN, G, D, H, W, tid); }
};
struct AggregateFwd {
    const float *corr, *weight; float* volume; int B, N, G, D, H, W;
    __device__ __forceinline__ void operator()(int64_t tid, int64_t) const { aggregate_fwd_thread(corr, weight, volume, B, N, G, D, H, W, tid); }
};
struct AggregateBwd {
    const float *gvol, *corr, *weight; float *gcorr, *gweight; int B, N, G, D, H, W;
    __device__ __forceinline__ void operator()(int64_t tid, int64_t) const { aggregate_bwd_thread(gvol, corr, weight, gcorr, gweight, B, N, G, D, H, W, tid); }
};
struct BnStats {
    const float* x; double* sums; int64_t M; int C;
    __device__ __forceinline__ void operator()(int64_t tid, int64_t nthreads) const { bn_stats_thread(x, sums, M, C, tid, nthreads); }
};
struct BnCollapse {
    double* sums; int C;
    __device__ __forceinline__ void operator()(int64_t tid, int64_t) const { bn_collapse_thread(sums, C, tid); }
};
struct BnFinalize {
    const double* sums; int replicas; double count; float eps, momentum; float *mean_invstd, *running_mean, *running_var; int C;
    __device__ __forceinline__ void operator()(int64_t tid, int64_t) const { bn_finalize_thread(sums, replicas, count, eps, momentum, mean_invstd, running_mean, running_var, C, tid); }
};
struct BnActFwd {
    const float *x, *mean_invstd, *gamma, *beta, *skip; float* y; int64_t M; int C, relu;
    __device__ __forceinline__ void operator()(int64_t tid, int64_t) const { bn_act_fwd_thread(x, mean_invstd, gamma, beta, skip, y, M, C, relu, tid); }
};
struct BnActBwdReduce {
    const float *gy, *x, *mean_invstd, *gamma, *beta; double* sums; int64_t M; int C, relu;
    __device__ __forceinline__ void operator()(int64_t tid, int64_t nthreads) const { bn_act_bwd_reduce_thread(gy, x, mean_invstd, gamma, beta, sums, M, C, relu, tid, nthreads); }
};
struct BnActBwdApply {
    const float *gy, *x, *mean_invstd, *gamma, *beta; const double* sums; double count; float* gx; int64_t M; int C, relu;
    __device__ __forceinline__ void operator()(int64_t tid, int64_t) const { bn_act_bwd_apply_thread(gy, x, mean_invstd, gamma, beta, sums, count, gx, M, C, relu, tid); }
};
template <int TS, int TB>
struct ConvWgrad {
    const float *small, *big; float* dw; WgradDims d;
    __device__ __forceinline__ void operator()(int64_t tid, int64_t) const { conv_wgrad_thread<TS, TB>(small, big, dw, d, tid); }
};
struct ThinConv {
    const float *x, *w, *bias; float* y; int B, D, H, W, Cin, Cout, kd, khw, act;
    __device__ __forceinline__ void operator()(int64_t tid, int64_t) const { thin_conv_thread(x, w, bias, y, B, D, H, W, Cin, Cout, kd, khw, act, tid); }
};
struct SigmoidBwd {
    const float *gy, *y; float* gx; int64_t n;
    __device__ __forceinline__ void operator()(int64_t tid, int64_t) const { sigmoid_bwd_thread(gy, y, gx, n, tid); }
};
struct HomoWarpBwd {
    const float *gwarped, *relproj, *depth; int depth_is_map; float* gsrc; int B, C, D, H, W;
    __device__ __forceinline__ void operator()(int64_t tid, int64_t) const { homo_warp_bwd_thread(gwarped, relproj, depth, depth_is_map, gsrc, B, C, D, H, W, tid); }
};
struct ProjMask {
    const float *relproj, *depth; float* mask; int B, N, D, H, W;
    __device__ __forceinline__ void operator()(int64_t tid, int64_t) const { proj_mask_thread(relproj, depth, mask, B, N, D, H, W, tid); }
};
struct EpipoleStats {
    const float *corr, *mask; float inv_t; float* stats; int BN, G, D; int64_t hw;
    __device__ __forceinline__ void operator()(int64_t tid, int64_t) const { epipole_stats_thread(corr, mask, inv_t, stats, BN, G, D, hw, tid); }
};
struct EpipoleAggregateFwd {
    const float *corr, *mask, *stats; float inv_t, inv_norm; float *volume, *wsum; int B, N, G, D; int64_t hw;
    __device__ __forceinline__ void operator()(int64_t tid, int64_t) const { epipole_aggregate_fwd_thread(corr, mask, stats, inv_t, inv_norm, volume, wsum, B, N, G, D, hw, tid); }
};
struct EpipoleAggregateBwd {
    const float *gvol, *corr, *mask, *stats, *volume, *wsum; float inv_t, inv_norm; float *gcorr, *gtemp; int B, N, G, D; int64_t hw;
    __device__ __forceinline__ void operator()(int64_t tid, int64_t) const { epipole_aggregate_bwd_thread(gvol, corr, mask, stats, volume, wsum, inv_t, inv_norm, gcorr, gtemp, B, N, G, D, hw, tid); }
};
struct HomoWarpBwdGrid {
    const float *gwarped, *src, *relproj, *depth; int depth_is_map; float *gdepth, *grelproj; int B, C, D, H, W;
    __device__ __forceinline__ void operator()(int64_t tid, int64_t) const { homo_warp_bwd_grid_thread(gwarped, src, relproj, depth, depth_is_map, gdepth, grelproj, B, C, D, H, W, tid); }
};
struct DepthRegressionBwd {
    const float *gdepth, *dv; int depth_is_map; float* gp; int B, D; int64_t hw;
    __device__ __forceinline__ void operator()(int64_t tid, int64_t) const { depth_regression_bwd_thread(gdepth, dv, depth_is_map, gp, B, D, hw, tid); }
};
struct MixupHead {
    const float *prob, *dv; float *depth, *conf; int B, D; int64_t hw;
    __device__ __forceinline__ void operator()(int64_t tid, int64_t) const { mixup_head_thread(prob, dv, depth, conf, B, D, hw, tid); }
};
struct SoftmaxBwd {
    const float *gp, *p; float* gpre; int B, D; int64_t hw;
    __device__ __forceinline__ void operator()(int64_t tid, int64_t) const { softmax_bwd_thread(gp, p, gpre, B, D, hw, tid); }
};

}  // namespace train
}  // namespace mvs
