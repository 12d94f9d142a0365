/*
 * mvs_b200.h — C ABI of libmvs_b200.so, the B200 (sm_100a) plane-sweep MVS depth engine.
 *
 * Drop-in boundary for ONE hot path of ewrfcas/MVSFormer: the per-reference-view cascade
 * (homography warp -> group-wise correlation cost volume -> 3D-CNN regularisation ->
 * temperature-softmax depth regression -> hypothesis re-scheduling).  The reference is pure
 * Python/PyTorch, so "the reference's FFI for this path" is its Python call surface; every
 * entry point below names the reference function (file:line in /root/reference) it replaces.
 * The Python mirror in mvsformer_b200/{warping,module,mvsformer_model}.py binds these with
 * ctypes and keeps the reference's names and signatures.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no torch types.  All tensor pointers are DEVICE
 *     pointers to fp32 unless marked [host].  The caller owns every buffer (inputs, outputs,
 *     workspaces); the library allocates nothing persistent.
 *   - `stream` is a cudaStream_t passed as void*; every call only enqueues work on it (no
 *     device synchronisation inside).
 *   - Return value: 0 on success, negative on error; mvs_last_error_string() gives the
 *     message of the last failing call on the calling thread.
 *   - "channels-last volume" = [B, D, H, W, C] with C innermost (the engine's internal layout
 *     for the cost volume and the 3D-CNN activations); "NCDHW" = PyTorch's default.
 */
#ifndef MVS_B200_H_
#define MVS_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MVS_OK 0
#define MVS_ERR_INVALID_ARGUMENT (-1)
#define MVS_ERR_UNSUPPORTED (-2)
#define MVS_ERR_CUDA (-3)

#define MVS_MAX_SRC_VIEWS 16

/* ---- library ------------------------------------------------------------------------------ */
int mvs_version(void);
const char* mvs_last_error_string(void);
/* Number of kernels this library has launched in this process (monotonic; used by bench.py). */
long long mvs_launch_count(void);
/* Fills SM count and compute capability of the current device. */
int mvs_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ---- A1. cameras: models/mvsformer_model.py:69-72 + models/warping.py:80-82 ----------------
 * proj_matrices [B,V,2,4,4] (extrinsic, intrinsic) -> relproj [B,V-1,12]: rows of the 3x4 matrix
 * [R|t] of  M = P_src * inverse(P_ref),  P = [K*E[:3,:4]; 0 0 0 1].  Evaluated in fp64 on the
 * device (one thread per source view), rounded once to fp32. */
int mvs_relative_projections(const float* proj_matrices, int B, int V, float* relproj, void* stream);
/* Same for already-composed 4x4 matrices (the homo_warping_* signature): src_proj, ref_proj [B,4,4]. */
int mvs_relative_projection_pair(const float* src_proj, const float* ref_proj, int B, float* relproj, void* stream);

/* ---- A2. homo_warping_3D / homo_warping_3D_with_mask: models/warping.py:69-109,155-189 ------
 * src_fea [B,C,H,W]; relproj [B,12]; depth [B,D] (depth_is_map=0) or [B,D,H,W] (1);
 * warped [B,C,D,H,W]; mask [B,D,H,W] uint8 (1 = out of bounds or z<=0) or NULL. */
int mvs_homo_warp(const float* src_fea, const float* relproj, const float* depth, int depth_is_map,
                  float* warped, uint8_t* mask, int B, int C, int D, int H, int W, void* stream);

/* ---- A3/A4/A6. cost-volume build: models/mvsformer_model.py:61-105,151-156 ------------------
 * Two sampling passes; the N x C x D x H x W warped tensor is never materialised.
 *   features: pointer to view 0 of batch 0 of a [B,V,C,H,W] tensor whose [C,H,W] blocks are
 *             contiguous; batch_stride / view_stride in elements.  View 0 is the reference view.
 *   depth    [B,D,H,W].
 * pass A: entropy [B,N,H,W] of softmax_d(sum_g corr) per source view (:88-90); optionally the
 *         eval-only cosine-similarity volume summed over views, sim_sum [B,D,H,W] (:81-85), NULL to skip.
 * pass B: volume [B,D,H,W,G] channels-last = sum_v w_v corr_v / (sum_v w_v + 1e-6) (:101-105);
 *         vis_weight [B,N,H,W]. */
int mvs_cost_volume_entropy(const float* features, int64_t batch_stride, int64_t view_stride,
                            const float* relproj, const float* depth, float* entropy, float* sim_sum,
                            int B, int V, int C, int G, int D, int H, int W, void* stream);
int mvs_cost_volume_aggregate(const float* features, int64_t batch_stride, int64_t view_stride,
                              const float* relproj, const float* depth, const float* vis_weight,
                              float* volume, int B, int V, int C, int G, int D, int H, int W, void* stream);
/* Same, with the volume rounded to TF32 on store (operand of the TF32 tensor-core 3D CNN). */
int mvs_cost_volume_aggregate_tf32(const float* features, int64_t batch_stride, int64_t view_stride,
                                   const float* relproj, const float* depth, const float* vis_weight,
                                   float* volume, int B, int V, int C, int G, int D, int H, int W, void* stream);
/* volume = sum_v w_v corr_v / (sum_v w_v + 1e-6) streamed over a stored per-view correlation corr [B,N,D,H,W,G]
 * (written by mvs_cost_volume_cl_entropy); models/mvsformer_model.py:97-105. */
int mvs_corr_aggregate(const float* corr, const float* vis_weight, float* volume, int B, int N, int D, int H, int W,
                       int round_tf32, void* stream);
/* Production cost-volume build over CHANNELS-LAST features (csrc/cost_volume_cl.cu): same arithmetic and outputs
 * as the entry points above (models/mvsformer_model.py:61-105, models/warping.py:84-107), features given as
 * feat_cl [B,V,H,W,C] (dense).  mvs_features_to_cl converts up to four NCHW tensors in ONE launch: segment s is
 * in[s] = [maps[s]][channels[s]][hw[s]]  ->  out[s] = [maps[s]][hw[s]][channels[s]]  (all five arrays are HOST arrays
 * of nseg entries; in/out entries are device pointers).
 * mvs_cost_volume_cl_entropy: pass A; sim_depth [B,H,W] (eval, may be NULL) receives depth[argmax_d of the cosine similarity
 * summed over views] directly (models/mvsformer_model.py:81-85,151-156) — the similarity volume never reaches HBM;
 * corr != NULL additionally stores the per-view group correlation [B,N,D,H,W,G]
 * (required where C/G >= 2, i.e. (C,D) = (64,32), (32,16), (16,8); must be NULL for (8,4), which keeps two sampling
 * passes so that the warped tensor never reaches HBM).  mvs_cost_volume_cl_aggregate: pass B for (C,D) = (8,4).
 * feat_cl holds nmaps maps [nmaps][H][W][C].  view_slots == NULL: the dense [B,V,H,W,C] tensor (nmaps = B*V).  view_slots !=
 * NULL (HOST array of V ints, B = 1): the reference view is map view_slots[0] and source view v is map view_slots[v] of a
 * per-scan feature pool — a reference view's views are addressed in place, nothing is gathered (StreamedCascade.run_scan).
 * Both return 1 (nothing launched) for shapes they do not cover; the caller then uses the NCHW entry points. */
int mvs_features_to_cl(const float* const* in, float* const* out, const int* channels, const int64_t* hw,
                       const int64_t* maps, int nseg, void* stream);
int mvs_cost_volume_cl_entropy(const float* feat_cl, int nmaps, const int* view_slots, const float* relproj, const float* depth,
                               float* entropy, float* sim_depth, float* corr, int B, int V, int C, int G, int D, int H, int W,
                               void* stream);
int mvs_cost_volume_cl_aggregate(const float* feat_cl, int nmaps, const int* view_slots, const float* relproj,
                                 const float* depth, const float* vis_weight, float* volume, int B, int V, int C, int G, int D,
                                 int H, int W, int round_tf32, void* stream);
/* sim_depth = depth[argmax_d sim_sum] (:151-156).  out [B,H,W]. */
int mvs_argmax_gather(const float* score, const float* depth, float* out, int B, int D, int H, int W, void* stream);

/* ---- A4. visibility net (StageNet.vis): models/mvsformer_model.py:37,91 ----------------------
 * entropy [M,H,W] -> weight [M,H,W]; three 3x3 conv (BN folded, ReLU) 1->16->16->8, 1x1 conv
 * 8->1 + bias, sigmoid.  params [host]: folded weights and shifts, packed as
 *   w1[16][9] b1[16] w2[16][16][9] b2[16] w3[8][16][9] b3[8] w4[8] b4[1]   (3641 floats). */
#define MVS_VIS_PARAM_FLOATS (16 * 9 + 16 + 16 * 16 * 9 + 16 + 8 * 16 * 9 + 8 + 8 + 1)
int mvs_vis_weight(const float* entropy, const float* params_host, float* weight, int M, int H, int W, void* stream);

/* ---- A7. 3D-CNN layers: models/module.py:83-159 (Conv3d / Deconv3d blocks), :469-594 --------
 * Channels-last activations.  y = act(conv(x) + shift) (+ skip);  BN (eval) is folded by the
 * caller: scale into the packed weights, shift passed per output channel.
 *   conv:   x [B,D,H,W,Cin] -> y [B,Do,Ho,Wo,Cout], kernel (kd,3,3) with kd in {1,3}, pad k/2, stride (sd,sh,sw) in {1,2};
 *           w packed [kd][3][3][Cin][Cout].
 *   deconv: ConvTranspose3d kernel (kd,3,3), pad k/2, output_padding = stride-1; x [B,D,H,W,Cin] -> y [B,D*sd,H*2,W*2,Cout];
 *           w packed [kd][3][3][Cin][Cout] (from torch's [Cin,Cout,kd,kh,kw]).
 *   skip (same shape as y) is added AFTER the activation (x = conv4 + conv7(x), module.py:500-502); NULL for none.
 *   shift NULL = zeros. */
int mvs_conv3d_cl(const float* x, const float* w_packed, const float* shift, const float* skip, float* y,
                  int B, int D, int H, int W, int Cin, int Cout, int kd, int sd, int sh, int sw, int relu, void* stream);
int mvs_deconv3d_cl(const float* x, const float* w_packed, const float* shift, const float* skip, float* y,
                    int B, int D, int H, int W, int Cin, int Cout, int kd, int sd, int relu, void* stream);
/* Tensor-core (tcgen05, kind::tf32, fp32 accumulate in TMEM) implicit-GEMM variant of mvs_conv3d_cl
 * for kernel (kd,3,3), stride sd in depth and shw (1 or 2) in both H and W.  Weights are packed by
 * the caller in operand order  [Cout_tiles][kd][3 kh][Cin/CS][3 kw][CS/4][n_tile][4]  with
 * CS = min(Cin, 32), rounded to TF32 (w_hi); w_lo = tf32(w - w_hi) selects the 3xTF32 mode
 * (fp32-grade accuracy), NULL selects plain TF32 (cuDNN's default conv math on GPUs). */
int mvs_conv3d_tc(const float* x, const float* w_hi, const float* w_lo, const float* shift, const float* skip,
                  float* y, int B, int D, int H, int W, int Cin, int Cout, int n_tile, int kd, int sd, int shw,
                  int relu, void* stream);
/* Tensor-core variant of mvs_deconv3d_cl (same geometry).  Weights packed in operand order
 * [Cout_tiles][kd][2 dy][Cin/CS][6 taps][CS/4][n_tile][4]; for dy = 0 the taps are
 * (kh=1,kw=0..2),(kh=2,kw=0..2), for dy = 1 (kh=0,kw=0..2) followed by three unused slots. */
int mvs_deconv3d_tc(const float* x, const float* w_hi, const float* w_lo, const float* shift, const float* skip,
                    float* y, int B, int D, int H, int W, int Cin, int Cout, int n_tile, int kd, int sd,
                    int relu, void* stream);
/* Diagnostic: nk MMAs (M=128, N, K=8) over caller-made shared-memory operand images with explicit
 * descriptor strides; dumps the 128 x N accumulator (used by tests to pin the operand layouts). */
int mvs_tc_probe(const float* a_img, int a_bytes, const float* b_img, int b_bytes, unsigned a_lbo,
                 unsigned a_sbo, unsigned b_lbo, unsigned b_sbo, int N, int nk, unsigned a_kstep,
                 unsigned b_kstep, float* d_out, void* stream);
/* Same with the A operand staged smem -> TMEM by tcgen05.cp.128x256b and read from tensor memory. */
int mvs_tc_probe_ts(const float* a_img, int a_bytes, const float* b_img, int b_bytes, unsigned a_lbo,
                    unsigned a_sbo, unsigned b_lbo, unsigned b_sbo, int N, int nk, unsigned a_kstep,
                    unsigned b_kstep, unsigned a_shift_bytes, float* d_out, void* stream);
/* Layout transforms at the module boundary (CostRegNet*.forward takes/returns NCDHW). */
int mvs_ncdhw_to_cl(const float* x, float* y, int B, int C, int D, int H, int W, void* stream);
int mvs_ncdhw_to_cl_tf32(const float* x, float* y, int B, int C, int D, int H, int W, void* stream);   /* + TF32 rounding */
int mvs_cl_to_ncdhw(const float* x, float* y, int B, int C, int D, int H, int W, void* stream);

/* ---- A7/A8. prob conv + head: models/module.py:493,582; models/mvsformer_model.py:110-125 ----
 * x channels-last [B,D,H,W,Cin] -> pre [B,D,H,W] with the 8->1 `prob` conv (ksize 1: 1x1x1 + bias;
 * ksize 3: 3x3x3, pad 1, no bias).  w [host]: [ksize^3][Cin], bias [host] 1 float or NULL. */
int mvs_prob_conv_cl(const float* x, const float* w_host, const float* bias_host, float* pre,
                     int B, int D, int H, int W, int Cin, int ksize, void* stream);
/* pre [B,D,H,W], depth_values [B,D,H,W] -> prob_volume = softmax_d(pre) [B,D,H,W] (NULL to skip),
 * photometric_confidence = max_d prob [B,H,W], depth [B,H,W]:
 *   mode 0 (eval):  sum_d softmax_d(pre * tmp) * depth_values   (module.py:597-603)
 *   mode 1 (train): depth_values[argmax_d prob]                 (mvsformer_model.py:117-120) */
int mvs_regression_head(const float* pre, const float* depth_values, float tmp, int mode,
                        float* prob_volume, float* depth, float* confidence, int B, int D, int H, int W, void* stream);
/* depth_regression(p, depth_values): module.py:597-603.  depth_is_map as in mvs_homo_warp. */
int mvs_depth_regression(const float* p, const float* depth_values, int depth_is_map, float* out,
                         int B, int D, int H, int W, void* stream);
/* conf_regression(p, n): module.py:606-619. */
int mvs_conf_regression(const float* p, int n, float* out, int B, int D, int H, int W, void* stream);

/* ---- A9. hypothesis schedules: models/module.py:622-699 -------------------------------------
 * cur_depth [B,ND] (the [B,192] range); out [B,D,H,W]. */
int mvs_init_inverse_range(const float* cur_depth, int ND, float* out, int B, int D, int H, int W, void* stream);
int mvs_init_range(const float* cur_depth, int ND, float* out, int B, int D, int H, int W, void* stream);
/* depth [B,H/2,W/2], depth_hypo [B,Dprev,H/2,W/2] -> out [B,D,H,W] (module.py:642-653). */
int mvs_schedule_inverse_range(const float* depth, const float* depth_hypo, int Dprev, float split_itv,
                               float* out, int B, int D, int H, int W, void* stream);
/* depth [B,H/2,W/2], interval [B] (device) -> out [B,D,H,W] (module.py:687-699). */
int mvs_schedule_range(const float* depth, const float* interval, float* out, int B, int D, int H, int W, void* stream);
/* Nearest-neighbour upsample + accumulate of the per-stage confidence (mvsformer_model.py:438-442):
 * acc[B,H,W] += scale * conf[B,h,w] (nearest, F.interpolate semantics). */
int mvs_confidence_accumulate(const float* conf, int h, int w, float* acc, int B, int H, int W, float scale, void* stream);
/* Same in one pass, also writing the nearest-upsampled stage confidence up[B,H,W] (the value the
 * reference stores back into outputs_stage['photometric_confidence'], mvsformer_model.py:439-441). */
int mvs_confidence_upsample_accumulate(const float* conf, int h, int w, float* up, float* acc, int B, int H, int W,
                                       float scale, void* stream);

/* ---- training path: train-mode forward + backward of StageNet (SURVEY.md 8b "Autograd") -------
 * The reference trains through torch autograd over models/mvsformer_model.py:61-125; the Python
 * mirror wraps the entry points below in torch.autograd.Function (mvsformer_b200/autograd.py).
 * In training the per-view correlation IS materialised (it is an autograd-saved tensor in the
 * reference too) as corr [B,N,D,H,W,G] channels-last; N = V-1 source views. */
/* corr_v = mean_c' ref * warped_v (mvsformer_model.py:70-79).  features/strides as mvs_cost_volume_entropy. */
int mvs_group_corr_fwd(const float* features, int64_t batch_stride, int64_t view_stride, const float* relproj,
                       const float* depth, float* corr, int B, int V, int C, int G, int D, int H, int W, void* stream);
/* Gradient w.r.t. the features of all views (the sampling grid is built under no_grad, warping.py:79):
 * gcorr [B,N,D,H,W,G] -> gfeat [B,V,C,H,W] dense, ZEROED by the caller, accumulated with fp32 atomics. */
int mvs_group_corr_bwd(const float* features, int64_t batch_stride, int64_t view_stride, const float* relproj,
                       const float* depth, const float* gcorr, float* gfeat, int B, int V, int C, int G, int D, int H,
                       int W, void* stream);
/* entropy [BN,H,W] of softmax_d(sum_g corr) (mvsformer_model.py:87-90; no gradient: the input is detached). */
int mvs_corr_entropy(const float* corr, float* entropy, int BN, int G, int D, int H, int W, void* stream);
/* volume [B,D,H,W,G] = sum_v corr_v w_v / (sum_v w_v + 1e-6), weight [B,N,H,W] (mvsformer_model.py:101-105), and its backward. */
int mvs_aggregate_fwd(const float* corr, const float* weight, float* volume, int B, int N, int G, int D, int H, int W,
                      void* stream);
int mvs_aggregate_bwd(const float* gvol, const float* corr, const float* weight, float* gcorr, float* gweight, int B, int N,
                      int G, int D, int H, int W, void* stream);
/* Train-mode BatchNorm over channels-last x [M,C] (nn.BatchNorm2d/3d inside ConvBnReLU / Conv3d / Deconv3d,
 * models/module.py:83-197).  sums: MVS_BN_REPLICAS x 2C doubles, ZEROED by the caller: the reduction kernels
 * spread their fp64 atomics over the replicas, mvs_bn_collapse folds them into the first 2C doubles (sum, sum of
 * squares), which is what finalize / apply read; with SyncBatchNorm the caller all-reduces those 2C doubles and
 * passes the global count.  mean_invstd: 2C floats.  running_* updated in place (NULL to skip):
 * r = (1-momentum) r + momentum * batch value, unbiased variance. */
#define MVS_BN_REPLICAS 32
int mvs_bn_stats(const float* x, double* sums, int64_t M, int C, void* stream);
int mvs_bn_collapse(double* sums, int C, void* stream);
/* replicas: MVS_BN_REPLICAS straight after mvs_bn_stats (finalize folds the copies itself), 1 after mvs_bn_collapse. */
int mvs_bn_finalize(const double* sums, int replicas, double count, float eps, float momentum, float* mean_invstd,
                    float* running_mean, float* running_var, int C, void* stream);
/* y = act((x - mean) * invstd * gamma + beta) (+ skip after the activation, module.py:500-502). */
int mvs_bn_act_fwd(const float* x, const float* mean_invstd, const float* gamma, const float* beta, const float* skip,
                   float* y, int64_t M, int C, int relu, void* stream);
/* Backward: sums (MVS_BN_REPLICAS x 2C doubles, ZEROED by the caller; mvs_bn_collapse afterwards) <- (d gamma, d beta); then
 * gx = gamma * invstd * (g - dbeta/count - xhat * dgamma/count), g = gy masked by the ReLU. */
int mvs_bn_act_bwd_reduce(const float* gy, const float* x, const float* mean_invstd, const float* gamma, const float* beta,
                          double* sums, int64_t M, int C, int relu, void* stream);
int mvs_bn_act_bwd_apply(const float* gy, const float* x, const float* mean_invstd, const float* gamma, const float* beta,
                         const double* sums, double count, float* gx, int64_t M, int C, int relu, void* stream);
/* Weight gradient of a (kd,k,k) conv / transposed conv (k in {1,3}, padding k/2), packed layout
 * dw [kd][k][k][Cin][Cout], ZEROED by the caller.  `small` lives on the strided grid, `big` on the fine
 * one (big position = small position * stride - pad + tap):
 *   conv   (small_is_cout = 1): small = grad of the output [.., Cout], big = the input [.., Cin]
 *   deconv (small_is_cout = 0): small = the input [.., Cin],  big = grad of the output [.., Cout] */
int mvs_conv_wgrad_cl(const float* small, const float* big, float* dw, int B, int Ds, int Hs, int Ws, int Db, int Hb,
                      int Wb, int Cs, int Cb, int kd, int khw, int sd, int shw, int small_is_cout, void* stream);
/* Few-channel convolution with DEVICE weights w [kd*k*k][Cin][Cout], bias [Cout] or NULL; stride 1, padding k/2;
 * act 0 none, 1 ReLU, 2 sigmoid (vis net 1->16 and 8->1, `prob` 8->1, and their data gradients). */
int mvs_thin_conv_cl(const float* x, const float* w, const float* bias, float* y, int B, int D, int H, int W, int Cin,
                     int Cout, int kd, int khw, int act, void* stream);
/* gx = gy * y * (1 - y) over n elements. */
int mvs_sigmoid_bwd(const float* gy, const float* y, float* gx, int64_t n, void* stream);
/* Backward of mvs_homo_warp w.r.t. src_fea (F.grid_sample's input gradient, warping.py:105): gwarped [B,C,D,H,W]
 * -> gsrc [B,C,H,W], ZEROED by the caller, fp32 atomics. */
int mvs_homo_warp_bwd(const float* gwarped, const float* relproj, const float* depth, int depth_is_map, float* gsrc,
                      int B, int C, int D, int H, int W, void* stream);
/* fusion_type 'epipole' / 'epipoleV2' (mvsformer_model.py:92-104): per-hypothesis softmax view weights.
 * corr [B,N,D,H,W,G] (mvs_group_corr_fwd), mask [B,N,D,H,W] 1/0 from mvs_proj_mask (V2) or NULL, weight
 * w_v[k] = softmax_k(sum_g corr_v / temperature - 10000 mask)[k] / norm -> volume [B,D,H,W,G], wsum [B,D,H,W],
 * stats [B,N,H,W,2] (softmax max / normaliser, kept for the backward).  Backward: gcorr [B,N,D,H,W,G] and the gradient
 * w.r.t. the temperature as 32 partial sums gtemp32 (ZEROED by the caller; NULL to skip). */
int mvs_proj_mask(const float* relproj, const float* depth, float* mask, int B, int N, int D, int H, int W, void* stream);
int mvs_epipole_aggregate_fwd(const float* corr, const float* mask, float temperature, float norm, float* stats,
                              float* volume, float* wsum, int B, int N, int G, int D, int H, int W, void* stream);
int mvs_epipole_aggregate_bwd(const float* gvol, const float* corr, const float* mask, const float* stats,
                              const float* volume, const float* wsum, float temperature, float norm, float* gcorr,
                              float* gtemp32, int B, int N, int G, int D, int H, int W, void* stream);
/* Backward of the warp through the sampling grid (diff_homo_warping_3D_with_mask, warping.py:112-152): gdepth [B,D,H,W] or
 * [B,D], grelproj [B, MVS_WARP_GRAD_REPLICAS, 12] (gradient of the rows of [R|t], spread over replicas: sum over axis 1);
 * both ZEROED by the caller. */
#define MVS_WARP_GRAD_REPLICAS 32
int mvs_homo_warp_bwd_grid(const float* gwarped, const float* src_fea, const float* relproj, const float* depth,
                           int depth_is_map, float* gdepth, float* grelproj, int B, int C, int D, int H, int W, void* stream);
/* Backward of mvs_depth_regression w.r.t. p (depth_type 're' in training): gp [B,D,H,W] = gdepth [B,H,W] * depth_values. */
int mvs_depth_regression_bwd(const float* gdepth, const float* depth_values, int depth_is_map, float* gp, int B, int D,
                             int H, int W, void* stream);
/* depth_type 'mixup_ce' head (models/mvsformer_model.py:126-136): prob, depth_values [B,D,H,W] -> depth, confidence [B,H,W]. */
int mvs_mixup_head(const float* prob, const float* depth_values, float* depth, float* confidence, int B, int D, int H, int W,
                   void* stream);
/* p = softmax_d(pre) on [B,D,H,W]: gpre = p * (gp - sum_d gp * p)  (mvsformer_model.py:111). */
int mvs_softmax_bwd(const float* gp, const float* p, float* gpre, int B, int D, int H, int W, void* stream);

/* ---- depth-map fusion: misc/fusion.py:69-118 as driven by test.py:404-435 -------------
 * Camera matrices are inverted by the caller in fp64 and passed per (batch, source view) as MVS_FUSION_MAT_FLOATS
 * floats: ref Kinv (9) Einv (16) E (16) K (9), then src Kinv (9) Einv (16) E (16) K (9).  All maps fp32; masks are
 * 1.0 / 0.0 like the reference's. */
#define MVS_FUSION_MAT_FLOATS 100
/* get_reproj (:79-98): ref_depth [n,h,w], src_depths [n,v,h,w] -> reproj_xyd [n,v,3,h,w], in_range [n,v,h,w]. */
int mvs_fusion_reproject(const float* ref_depth, const float* src_depths, const float* mats, float* reproj_xyd,
                         float* in_range, int N, int V, int H, int W, void* stream);
/* vis_filter (:101-109) + ave_fusion (:112-114): -> masks [n,v,h,w], mask [n,h,w], ave [n,h,w]. */
int mvs_fusion_filter(const float* ref_depth, const float* reproj_xyd, const float* in_range, float img_dist_thresh,
                      float depth_thresh, float vthresh, float* masks, float* mask, float* ave, int N, int V, int H, int W,
                      void* stream);
/* get_reproj_dynamic (:116-152) and vis_filter_dynamic (:155-168) + the vote / averaging of test.py:502-511:
 * vis_mask [n,v,h,w] (level k = v), geo_mask [n,h,w], ave [n,h,w], level_counts [n,v-1,h,w] or NULL. */
int mvs_fusion_reproject_dynamic(const float* ref_depth, const float* src_depths, const float* mats, float* reproj_xyd,
                                 int N, int V, int H, int W, void* stream);
int mvs_fusion_filter_dynamic(const float* ref_depth, const float* reproj_xyd, float dist_base, float rel_diff_base,
                              float* vis_mask, float* geo_mask, float* ave, float* level_counts, int N, int V, int H, int W,
                              void* stream);
/* World points of a depth map (test.py:433-435): depth [n,h,w], mats [n,25] = Kinv (9) Einv (16) -> points [n,3,h,w]. */
int mvs_fusion_points(const float* depth, const float* mats, float* points, int N, int H, int W, void* stream);
/* prob_filter (:69-76): prob [n,c,h,w], thresholds [host] (<= 8) -> mask [n,h,w] = AND_i prob[:, i] > thresh[i]. */
int mvs_fusion_prob_filter(const float* prob, const float* thresh_host, int nthresh, float* mask, int N, int C, int H, int W,
                           void* stream);

/* ---- fused visibility net (csrc/vis_fused.cu): models/mvsformer_model.py:37,91 in ONE kernel --------------------------
 * entropy [M,H,W] -> weight [M,H,W]; only those two maps touch HBM.  params [host]: w1[16][9] b1[16] shift2[16] shift3[8]
 * w4[8] b4 (BN folded into w1 / the packed weights; shifts = folded BN biases), MVS_VIS_FUSED_PARAM_FLOATS floats.
 * w2 (16->16) and w3 (16->8): device, TF32, packed [kw][4 input-channel quads][rows][4] where the rows of a (kw, quad) are
 * [kh][cout] — 48 rows for w2, 24 + 8 rows of zeros for w3 (mvsformer_b200.engine.pack_vis_fused_weights): the three kernel
 * rows are ONE tcgen05.mma operand. */
#define MVS_VIS_FUSED_PARAM_FLOATS (16 * 9 + 16 + 16 + 8 + 8 + 1)
int mvs_vis_fused(const float* entropy, const float* params, const float* w2, const float* w3, float* weight, int M, int H,
                  int W, void* stream);

/* ---- round-2 persistent TMA-fed tcgen05 convolutions (csrc/conv3d_tma.cu) ---------------------------------------------
 * Depth-unstrided layers of CostRegNet3D (models/module.py:550-594) and the 3x3 tensor-core layers of the visibility net
 * (models/mvsformer_model.py:37).  x [B,D,H,W,Cin] channels-last with TF32-rounded values, y [B,D,H,W,Cout]
 * (TF32-rounded), shift [Cout] or NULL (folded BN / bias), skip like y or NULL, relu applied before the skip add.
 * w is packed by the host (mvsformer_b200.engine.pack_tma_weights): [Cout tiles][kh][kw][Cin/4][kd][n_tile][4].
 * Kernel (kd,3,3), kd in {1,3}, depth stride 1, padding (kd/2,1,1).  mode 0: stride 1, y [B,D,H,W,Cout]; mode 1: stride
 * (1,2,2), y [B,D,ceil(H/2),ceil(W/2),Cout]; mode 2: transposed convolution, stride (1,2,2), output_padding (0,1,1),
 * y [B,D,2H,2W,Cout] (w from torch's ConvTranspose3d weight permuted to [kd,kh,kw,Cin,Cout] before packing).
 * Accumulators live in tensor memory: (mode 2 ? 4 : 1) * D * n_tile <= 512 columns (double-buffered when twice that fits). */
int mvs_conv3d_tma(const float* x, const float* w, const float* shift, const float* skip, float* y, int B, int D, int H,
                   int W, int Cin, int Cout, int n_tile, int kd, int mode, int relu, void* stream);

/* The regulariser's last two layers in one kernel (CostRegNet3D eval, models/module.py:575,582,592-593): mode-2 transposed
 * convolution 16 -> 8 (+ shift, ReLU, + skip) with the 1x1x1 `prob` convolution 8 -> 1 (+ bias) applied to its result in the
 * epilogue; the 8-channel tensor never reaches HBM.  x [B,D,H,W,16], w packed as for mvs_conv3d_tma (mode 2, n_tile 16),
 * skip [B,D,2H,2W,8] or NULL, prob_w_host [8] (HOST), pre [B,D,2H,2W].  Bit-identical to mvs_conv3d_tma followed by
 * mvs_prob_conv_cl (ksize 1). */
int mvs_conv3d_tma_prob(const float* x, const float* w, const float* shift, const float* skip, const float* prob_w_host,
                        float prob_bias, float* pre, int B, int D, int H, int W, int Cin, int kd, int relu, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MVS_B200_H_ */
