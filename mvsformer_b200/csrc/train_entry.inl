// train_entry.inl — C-ABI entry points of the training path (declared in include/mvs_b200.h,
// section "training path").  Included by train.cu (CUDA build: launch_flat launches a
// __global__ wrapper on the caller's stream) and by tests/emu/emu.cpp (CPU emulation used only
// by the test-suite: launch_flat loops over the thread ids on host memory).
//
// The includer provides MVS_REQUIRE, MVS_OK, train_kernels.cuh and
//   template <class F> int launch_flat(const F& f, int64_t nthreads, void* stream, const char* name);
// which runs f(tid, T) for tid in [0, T), T = nthreads rounded up to a multiple of 256.

namespace mvs {
namespace train {

static inline bool pow2_le_256(int c) { return c >= 1 && c <= 256 && (c & (c - 1)) == 0; }

// threads for a grid-stride reduction over `elems` elements: enough to fill the GPU, a
// multiple of 256 (hence of every supported channel count)
static inline int64_t reduce_threads(int64_t elems) {
    int64_t t = (elems + 15) / 16;
    const int64_t cap = 148 * 8 * 256;
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

extern "C" int mvs_bn_finalize(const double* sums, double count, float eps, float momentum, float* mean_invstd,
                               float* running_mean, float* running_var, int C, void* stream) {
    using namespace mvs::train;
    MVS_REQUIRE(sums && mean_invstd, "mvs_bn_finalize: null pointer");
    MVS_REQUIRE(count >= 1.0 && C >= 1, "mvs_bn_finalize: bad count / C");
    BnFinalize f{sums, count, eps, momentum, mean_invstd, running_mean, running_var, C};
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
    WgradDims d{B, Ds, Hs, Ws, Db, Hb, Wb, Cs, Cb, kd, khw, sd, shw, small_is_cout ? 1 : 0};
    const int ts = Cs % 4 == 0 ? 4 : 1, tb = Cb % 4 == 0 ? 4 : 1;
    const int64_t threads = (int64_t)kd * khw * khw * (Cs / ts) * (Cb / tb) * B * Ds * Hs;
    if (ts == 4 && tb == 4) return launch_flat(ConvWgrad<4, 4>{small, big, dw, d}, threads, stream, "conv_wgrad<4,4>");
    if (ts == 4) return launch_flat(ConvWgrad<4, 1>{small, big, dw, d}, threads, stream, "conv_wgrad<4,1>");
    if (tb == 4) return launch_flat(ConvWgrad<1, 4>{small, big, dw, d}, threads, stream, "conv_wgrad<1,4>");
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
