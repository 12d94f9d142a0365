// train_entry.inl — C-ABI entry points of the training path (declared in include/mvs_b200.h,
// section "training path").  Included by train.cu (CUDA build: launch_flat launches a
// __global__ wrapper on the caller's stream) and by tests/emu/emu.cpp (CPU emulation used only
// by the test-suite: launch_flat loops over the thread ids on host memory).
//
// The includer provides MVS_REQUIRE, MVS_OK, train_kernels.cuh and
//   template <class F> int launch_flat(const F& f, int64_t nthreads, void* stream, const char* name);
// which runs f(tid, T) for tid in [0, T), T = nthreads rounded up to a multiple of 256.

#include <stdlib.h>

namespace mvs {
namespace train {

static inline bool pow2_le_256(int c) { return c >= 1 && c <= 256 && (c & (c - 1)) == 0; }

// threads for a grid-stride reduction over `elems` elements: enough to fill the GPU, a
// multiple of 256 (hence of every supported channel count)
static inline int64_t reduce_threads(int64_t elems) {
    int64_t t = (elems + 15) / 16;
    const int64_t cap = 148 * 4 * 256;
    if (t > cap) t = cap;
    if (t < 256) t = 256;
    return (t + 255) / 256 * 256;
}

static int check_corr(const char* fn, const void* a, const void* b, const void* c, const void* d, int B, int V, int C, int G,
                      int D, int H, int W) {
    MVS_REQUIRE(a && b && c && d, "%s: null pointer", fn);
    MVS_REQUIRE(B >= 1 && V >= 2 && D >= 1 && H >= 1 && W >= 1, "%s: empty shape B=%d V=%d D=%d H=%d W=%d", fn, B, V, D, H, W);
    MVS_REQUIRE(G >= 1 && C >= G && C % G == 0, "%s: C=%d is not a multiple of G=%d", fn, C, G);
    MVS_REQUIRE(V - 1 <= MVS_TRAIN_MAX_VIEWS, "%s: at most %d source views", fn, MVS_TRAIN_MAX_VIEWS);
    return MVS_OK;
}

}  // namespace train
}  // namespace mvs

extern "C" int mvs_group_corr_fwd(const float* features, int64_t batch_stride, int64_t view_stride, const float* relproj,
                                  const float* depth, float* corr, int B, int V, int C, int G, int D, int H, int W,
                                  void* stream) {
    using namespace mvs::train;
    int rc = check_corr("mvs_group_corr_fwd", features, relproj, depth, corr, B, V, C, G, D, H, W);
    if (rc) return rc;
    GroupCorrFwd f{features, relproj, depth, corr, CorrDims{B, V, C, G, D, H, W, batch_stride, view_stride}};
    return launch_flat(f, (int64_t)B * (V - 1) * D * H * W, stream, "group_corr_fwd");
}

extern "C" int mvs_group_corr_bwd(const float* features, int64_t batch_stride, int64_t view_stride, const float* relproj,
                                  const float* depth, const float* gcorr, float* gfeat, int B, int V, int C, int G, int D,
                                  int H, int W, void* stream) {
    using namespace mvs::train;
    int rc = check_corr("mvs_group_corr_bwd", features, relproj, depth, gcorr, B, V, C, G, D, H, W);
    if (rc) return rc;
    MVS_REQUIRE(gfeat, "mvs_group_corr_bwd: null gradient output");
    GroupCorrBwd f{features, relproj, depth, gcorr, gfeat, CorrDims{B, V, C, G, D, H, W, batch_stride, view_stride}};
    return launch_flat(f, (int64_t)B * (V - 1) * D * H * W, stream, "group_corr_bwd");
}

extern "C" int mvs_corr_entropy(const float* corr, float* entropy, int BN, int G, int D, int H, int W, void* stream) {
    using namespace mvs::train;
    MVS_REQUIRE(corr && entropy, "mvs_corr_entropy: null pointer");
    MVS_REQUIRE(BN >= 1 && G >= 1 && D >= 1 && H >= 1 && W >= 1, "mvs_corr_entropy: empty shape");
    CorrEntropy f{corr, entropy, BN, G, D, H, W};
    return launch_flat(f, (int64_t)BN * H * W, stream, "corr_entropy");
}

extern "C" int mvs_aggregate_fwd(const float* corr, const float* weight, float* volume, int B, int N, int G, int D, int H,
                                 int W, void* stream) {
    using namespace mvs::train;
    MVS_REQUIRE(corr && weight && volume, "mvs_aggregate_fwd: null pointer");
    MVS_REQUIRE(B >= 1 && N >= 1 && G >= 1 && D >= 1 && H >= 1 && W >= 1, "mvs_aggregate_fwd: empty shape");
    AggregateFwd f{corr, weight, volume, B, N, G, D, H, W};
    return launch_flat(f, (int64_t)B * D * H * W * G, stream, "aggregate_fwd");
}

extern "C" int mvs_aggregate_bwd(const float* gvol, const float* corr, const float* weight, float* gcorr, float* gweight,
                                 int B, int N, int G, int D, int H, int W, void* stream) {
    using namespace mvs::train;
    MVS_REQUIRE(gvol && corr && weight && gcorr && gweight, "mvs_aggregate_bwd: null pointer");
    MVS_REQUIRE(B >= 1 && N >= 1 && G >= 1 && D >= 1 && H >= 1 && W >= 1, "mvs_aggregate_bwd: empty shape");
    MVS_REQUIRE(N <= MVS_TRAIN_MAX_VIEWS, "mvs_aggregate_bwd: at most %d source views", MVS_TRAIN_MAX_VIEWS);
    AggregateBwd f{gvol, corr, weight, gcorr, gweight, B, N, G, D, H, W};
    return launch_flat(f, (int64_t)B * H * W, stream, "aggregate_bwd");
}

extern "C" int mvs_bn_stats(const float* x, double* sums, int64_t M, int C, void* stream) {
    using namespace mvs::train;
    MVS_REQUIRE(x && sums, "mvs_bn_stats: null pointer");
    MVS_REQUIRE(M >= 1 && pow2_le_256(C), "mvs_bn_stats: need M >= 1 and C a power of two <= 256 (M=%lld C=%d)", (long long)M, C);
    BnStats f{x, sums, M, C};
    return launch_flat(f, reduce_threads(M * C), stream, "bn_stats");
}

extern "C" int mvs_bn_collapse(double* sums, int C, void* stream) {
    using namespace mvs::train;
    MVS_REQUIRE(sums && C >= 1, "mvs_bn_collapse: null pointer or C < 1");
    BnCollapse f{sums, C};
    return launch_flat(f, 2 * C, stream, "bn_collapse");
}

extern "C" int mvs_bn_finalize(const double* sums, int replicas, double count, float eps, float momentum,
                               float* mean_invstd, float* running_mean, float* running_var, int C, void* stream) {
    using namespace mvs::train;
    MVS_REQUIRE(sums && mean_invstd, "mvs_bn_finalize: null pointer");
    MVS_REQUIRE(count >= 1.0 && C >= 1, "mvs_bn_finalize: bad count / C");
    MVS_REQUIRE(replicas == 1 || replicas == MVS_BN_REPLICAS, "mvs_bn_finalize: replicas must be 1 or %d", MVS_BN_REPLICAS);
    BnFinalize f{sums, replicas, count, eps, momentum, mean_invstd, running_mean, running_var, C};
    return launch_flat(f, C, stream, "bn_finalize");
}

extern "C" int mvs_bn_act_fwd(const float* x, const float* mean_invstd, const float* gamma, const float* beta,
                              const float* skip, float* y, int64_t M, int C, int relu, void* stream) {
    using namespace mvs::train;
    MVS_REQUIRE(x && mean_invstd && gamma && beta && y, "mvs_bn_act_fwd: null pointer");
    MVS_REQUIRE(M >= 1 && C >= 1, "mvs_bn_act_fwd: empty shape");
    BnActFwd f{x, mean_invstd, gamma, beta, skip, y, M, C, relu};
    return launch_flat(f, M * C, stream, "bn_act_fwd");
}

extern "C" int mvs_bn_act_bwd_reduce(const float* gy, const float* x, const float* mean_invstd, const float* gamma,
                                     const float* beta, double* sums, int64_t M, int C, int relu, void* stream) {
    using namespace mvs::train;
    MVS_REQUIRE(gy && x && mean_invstd && gamma && beta && sums, "mvs_bn_act_bwd_reduce: null pointer");
    MVS_REQUIRE(M >= 1 && pow2_le_256(C), "mvs_bn_act_bwd_reduce: need M >= 1 and C a power of two <= 256");
    BnActBwdReduce f{gy, x, mean_invstd, gamma, beta, sums, M, C, relu};
    return launch_flat(f, reduce_threads(M * C), stream, "bn_act_bwd_reduce");
}

extern "C" int mvs_bn_act_bwd_apply(const float* gy, const float* x, const float* mean_invstd, const float* gamma,
                                    const float* beta, const double* sums, double count, float* gx, int64_t M, int C,
                                    int relu, void* stream) {
    using namespace mvs::train;
    MVS_REQUIRE(gy && x && mean_invstd && gamma && beta && sums && gx, "mvs_bn_act_bwd_apply: null pointer");
    MVS_REQUIRE(M >= 1 && C >= 1 && count >= 1.0, "mvs_bn_act_bwd_apply: empty shape");
    BnActBwdApply f{gy, x, mean_invstd, gamma, beta, sums, count, gx, M, C, relu};
    return launch_flat(f, M * C, stream, "bn_act_bwd_apply");
}

extern "C" int mvs_conv_wgrad_cl(const float* small, const float* big, float* dw, int B, int Ds, int Hs, int Ws, int Db,
                                 int Hb, int Wb, int Cs, int Cb, int kd, int khw, int sd, int shw, int small_is_cout,
                                 void* stream) {
    using namespace mvs::train;
    MVS_REQUIRE(small && big && dw, "mvs_conv_wgrad_cl: null pointer");
    MVS_REQUIRE(B >= 1 && Ds >= 1 && Hs >= 1 && Ws >= 1 && Db >= 1 && Hb >= 1 && Wb >= 1 && Cs >= 1 && Cb >= 1,
                "mvs_conv_wgrad_cl: empty shape");
    MVS_REQUIRE((kd == 1 || kd == 3) && (khw == 1 || khw == 3), "mvs_conv_wgrad_cl: kernel sizes must be 1 or 3");
    MVS_REQUIRE((sd == 1 || sd == 2) && (shw == 1 || shw == 2), "mvs_conv_wgrad_cl: strides must be 1 or 2");
    // the fine grid must be the one a (transposed) convolution of this geometry produces
    MVS_REQUIRE((Db + sd - 1) / sd == Ds && (Hb + shw - 1) / shw == Hs && (Wb + shw - 1) / shw == Ws,
                "mvs_conv_wgrad_cl: grids [%d,%d,%d] and [%d,%d,%d] do not match strides (%d,%d,%d)", Ds, Hs, Ws, Db, Hb,
                Wb, sd, shw, shw);
    // register tile of a thread: 8x8 channels (four 128-bit loads per 64 FMAs) when both channel counts allow,
    // else 4 or 1 per side.  The first version (4x4 tile, scalar loads) was load-instruction bound:
    // 13 of 40 ms of a cfg-5 training step (profiles/r01_train_step_entry_points_final.json).
    // Measured on B200 (profiles/r02_ab_variants.json): the 8x8 tile is SLOWER (36.2 vs 13.6 ms of weight gradients per
    // cfg-5 step: a quarter of the threads, 152 registers), so 4 per side is the default; MVS_WGRAD_TILE=8 selects it.
    int tmax = 4;
    if (const char* env = getenv("MVS_WGRAD_TILE")) tmax = atoi(env);
    const int ts = (Cs % 8 == 0 && tmax >= 8) ? 8 : (Cs % 4 == 0 && tmax >= 4 ? 4 : 1);
    const int tb = (Cb % 8 == 0 && tmax >= 8) ? 8 : (Cb % 4 == 0 && tmax >= 4 ? 4 : 1);
    MVS_REQUIRE((ts == 1 || ((uintptr_t)small & 15) == 0) && (tb == 1 || ((uintptr_t)big & 15) == 0),
                "mvs_conv_wgrad_cl: tensors with a multiple of 4 channels must be 16-byte aligned");
    const int64_t ntasks = (int64_t)kd * khw * khw * (Cs / ts) * (Cb / tb), nrows = (int64_t)B * Ds * Hs;
    // Work split.  A thread owns (tap, channel tile) x a unit of voxels and ends with ts*tb atomics.  Aim at
    // ~512k threads: layers with many (row, task) pairs take several rows per thread (fewer atomics), layers
    // with few split each row into x segments (more threads; the serial loop over x is the latency chain).
    // MVS_WGRAD_ROWS / MVS_WGRAD_XSEG override (tests).
    const int64_t target = 512 * 1024;
    int64_t rpt = ntasks * nrows / target, xseg = Ws;
    if (rpt < 1) {
        rpt = 1;
        const int64_t want_seg = target / (ntasks * nrows);          // >= 1 here
        xseg = (Ws + want_seg - 1) / want_seg;
        if (xseg < 16) xseg = 16;
    }
    if (const char* env = getenv("MVS_WGRAD_ROWS")) rpt = atoi(env);
    if (const char* env = getenv("MVS_WGRAD_XSEG")) xseg = atoi(env);
    if (rpt < 1) rpt = 1;
    if (rpt > 32) rpt = 32;
    if (rpt > nrows) rpt = nrows;
    if (xseg < 1 || xseg > Ws) xseg = Ws;
    WgradDims d{B, Ds, Hs, Ws, Db, Hb, Wb, Cs, Cb, kd, khw, sd, shw, small_is_cout ? 1 : 0, (int)rpt, (int)xseg};
    const int64_t threads = ntasks * ((nrows + rpt - 1) / rpt) * ((Ws + xseg - 1) / xseg);
#define MVS_WGRAD_CASE(A, Bt) \
    if (ts == A && tb == Bt) return launch_flat(ConvWgrad<A, Bt>{small, big, dw, d}, threads, stream, "conv_wgrad<" #A "," #Bt ">")
    MVS_WGRAD_CASE(8, 8); MVS_WGRAD_CASE(8, 4); MVS_WGRAD_CASE(4, 8); MVS_WGRAD_CASE(4, 4);
    MVS_WGRAD_CASE(8, 1); MVS_WGRAD_CASE(1, 8); MVS_WGRAD_CASE(4, 1); MVS_WGRAD_CASE(1, 4);
#undef MVS_WGRAD_CASE
    return launch_flat(ConvWgrad<1, 1>{small, big, dw, d}, threads, stream, "conv_wgrad<1,1>");
}

extern "C" int mvs_thin_conv_cl(const float* x, const float* w, const float* bias, float* y, int B, int D, int H, int W,
                                int Cin, int Cout, int kd, int khw, int act, void* stream) {
    using namespace mvs::train;
    MVS_REQUIRE(x && w && y, "mvs_thin_conv_cl: null pointer");
    MVS_REQUIRE(B >= 1 && D >= 1 && H >= 1 && W >= 1 && Cin >= 1 && Cout >= 1, "mvs_thin_conv_cl: empty shape");
    MVS_REQUIRE((kd == 1 || kd == 3) && (khw == 1 || khw == 3), "mvs_thin_conv_cl: kernel sizes must be 1 or 3");
    MVS_REQUIRE(act >= 0 && act <= 2, "mvs_thin_conv_cl: act must be 0 (none), 1 (ReLU) or 2 (sigmoid)");
    ThinConv f{x, w, bias, y, B, D, H, W, Cin, Cout, kd, khw, act};
    return launch_flat(f, (int64_t)B * D * H * W * Cout, stream, "thin_conv");
}

extern "C" int mvs_sigmoid_bwd(const float* gy, const float* y, float* gx, int64_t n, void* stream) {
    using namespace mvs::train;
    MVS_REQUIRE(gy && y && gx && n >= 1, "mvs_sigmoid_bwd: null pointer or empty");
    SigmoidBwd f{gy, y, gx, n};
    return launch_flat(f, n, stream, "sigmoid_bwd");
}

extern "C" int mvs_softmax_bwd(const float* gp, const float* p, float* gpre, int B, int D, int H, int W, void* stream) {
    using namespace mvs::train;
    MVS_REQUIRE(gp && p && gpre, "mvs_softmax_bwd: null pointer");
    MVS_REQUIRE(B >= 1 && D >= 1 && H >= 1 && W >= 1, "mvs_softmax_bwd: empty shape");
    SoftmaxBwd f{gp, p, gpre, B, D, (int64_t)H * W};
    return launch_flat(f, (int64_t)B * H * W, stream, "softmax_bwd");
}

extern "C" int mvs_homo_warp_bwd(const float* gwarped, const float* relproj, const float* depth, int depth_is_map,
                                 float* gsrc, int B, int C, int D, int H, int W, void* stream) {
    using namespace mvs::train;
    MVS_REQUIRE(gwarped && relproj && depth && gsrc, "mvs_homo_warp_bwd: null pointer");
    MVS_REQUIRE(B >= 1 && C >= 1 && D >= 1 && H >= 1 && W >= 1, "mvs_homo_warp_bwd: empty shape");
    HomoWarpBwd f{gwarped, relproj, depth, depth_is_map, gsrc, B, C, D, H, W};
    return launch_flat(f, (int64_t)B * D * H * W, stream, "homo_warp_bwd");
}

extern "C" int mvs_depth_regression_bwd(const float* gdepth, const float* depth_values, int depth_is_map, float* gp, int B,
                                        int D, int H, int W, void* stream) {
    using namespace mvs::train;
    MVS_REQUIRE(gdepth && depth_values && gp, "mvs_depth_regression_bwd: null pointer");
    MVS_REQUIRE(B >= 1 && D >= 1 && H >= 1 && W >= 1, "mvs_depth_regression_bwd: empty shape");
    DepthRegressionBwd f{gdepth, depth_values, depth_is_map, gp, B, D, (int64_t)H * W};
    return launch_flat(f, (int64_t)B * D * H * W, stream, "depth_regression_bwd");
}

extern "C" int mvs_mixup_head(const float* prob, const float* depth_values, float* depth, float* confidence, int B, int D,
                              int H, int W, void* stream) {
    using namespace mvs::train;
    MVS_REQUIRE(prob && depth_values && depth && confidence, "mvs_mixup_head: null pointer");
    MVS_REQUIRE(B >= 1 && D >= 2 && H >= 1 && W >= 1, "mvs_mixup_head: need at least two hypotheses");
    MixupHead f{prob, depth_values, depth, confidence, B, D, (int64_t)H * W};
    return launch_flat(f, (int64_t)B * H * W, stream, "mixup_head");
}

extern "C" int mvs_homo_warp_bwd_grid(const float* gwarped, const float* src_fea, const float* relproj, const float* depth,
                                      int depth_is_map, float* gdepth, float* grelproj, int B, int C, int D, int H, int W,
                                      void* stream) {
    using namespace mvs::train;
    MVS_REQUIRE(gwarped && src_fea && relproj && depth && gdepth && grelproj, "mvs_homo_warp_bwd_grid: null pointer");
    MVS_REQUIRE(B >= 1 && C >= 1 && D >= 1 && H >= 1 && W >= 1, "mvs_homo_warp_bwd_grid: empty shape");
    HomoWarpBwdGrid f{gwarped, src_fea, relproj, depth, depth_is_map, gdepth, grelproj, B, C, D, H, W};
    return launch_flat(f, (int64_t)B * D * H * W, stream, "homo_warp_bwd_grid");
}

// ---- fusion_type 'epipole' / 'epipoleV2' (models/mvsformer_model.py:92-104) ------------------------------------------
extern "C" int mvs_proj_mask(const float* relproj, const float* depth, float* mask, int B, int N, int D, int H, int W,
                             void* stream) {
    using namespace mvs::train;
    MVS_REQUIRE(relproj && depth && mask, "mvs_proj_mask: null pointer");
    MVS_REQUIRE(B >= 1 && N >= 1 && D >= 1 && H >= 1 && W >= 1, "mvs_proj_mask: empty shape");
    ProjMask f{relproj, depth, mask, B, N, D, H, W};
    return launch_flat(f, (int64_t)B * N * D * H * W, stream, "proj_mask");
}

extern "C" int mvs_epipole_aggregate_fwd(const float* corr, const float* mask, float temperature, float norm, float* stats,
                                         float* volume, float* wsum, int B, int N, int G, int D, int H, int W, void* stream) {
    using namespace mvs::train;
    MVS_REQUIRE(corr && stats && volume && wsum, "mvs_epipole_aggregate_fwd: null pointer");
    MVS_REQUIRE(B >= 1 && N >= 1 && G >= 1 && G <= 8 && D >= 1 && H >= 1 && W >= 1, "mvs_epipole_aggregate_fwd: bad shape (G <= 8)");
    MVS_REQUIRE(temperature > 0.0f && norm > 0.0f, "mvs_epipole_aggregate_fwd: temperature and norm must be positive");
    const int64_t hw = (int64_t)H * W;
    EpipoleStats f1{corr, mask, 1.0f / temperature, stats, B * N, G, D, hw};
    int rc = launch_flat(f1, (int64_t)B * N * hw, stream, "epipole_stats");
    if (rc) return rc;
    EpipoleAggregateFwd f2{corr, mask, stats, 1.0f / temperature, 1.0f / norm, volume, wsum, B, N, G, D, hw};
    return launch_flat(f2, (int64_t)B * D * hw, stream, "epipole_aggregate_fwd");
}

extern "C" int mvs_epipole_aggregate_bwd(const float* gvol, const float* corr, const float* mask, const float* stats,
                                         const float* volume, const float* wsum, float temperature, float norm, float* gcorr,
                                         float* gtemp32, int B, int N, int G, int D, int H, int W, void* stream) {
    using namespace mvs::train;
    MVS_REQUIRE(gvol && corr && stats && volume && wsum && gcorr, "mvs_epipole_aggregate_bwd: null pointer");
    MVS_REQUIRE(B >= 1 && N >= 1 && G >= 1 && G <= 8 && D >= 1 && H >= 1 && W >= 1, "mvs_epipole_aggregate_bwd: bad shape (G <= 8)");
    MVS_REQUIRE(temperature > 0.0f && norm > 0.0f, "mvs_epipole_aggregate_bwd: temperature and norm must be positive");
    EpipoleAggregateBwd f{gvol, corr, mask, stats, volume, wsum, 1.0f / temperature, 1.0f / norm, gcorr, gtemp32, B, N, G, D,
                          (int64_t)H * W};
    return launch_flat(f, (int64_t)B * N * H * W, stream, "epipole_aggregate_bwd");
}
