"""Thin tensor-level wrappers over the C ABI (include/mvs_b200.h).

Each function allocates its outputs with torch's caching allocator on the current device,
enqueues the kernel(s) on torch's current CUDA stream and returns immediately.  PyTorch is only
the owner of device memory and streams here; all arithmetic is in libmvs_b200.so.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import check, ptr, require_cuda, stream


def require_cuda_device(t):
    """Device check for tensors whose strides are free (the feature views): CUDA only, there is no CPU path."""
    if not t.is_cuda:
        raise RuntimeError("mvsformer_b200 runs on CUDA tensors only; there is no CPU path")
    _lib.check_device(t)


def _f32(t):
    return t if t.dtype == torch.float32 else t.float()


def relative_projections(proj_matrices):
    """[B,V,2,4,4] -> relproj [B,V-1,12]  (mvsformer_model.py:69-72 + warping.py:80-82)."""
    proj_matrices = _f32(proj_matrices).contiguous()
    require_cuda(proj_matrices)
    if proj_matrices.dim() != 5 or tuple(proj_matrices.shape[2:]) != (2, 4, 4):
        raise RuntimeError("proj_matrices must be [B,V,2,4,4], got %s" % (tuple(proj_matrices.shape),))
    b, v = proj_matrices.shape[:2]
    out = torch.empty(b, v - 1, 12, device=proj_matrices.device, dtype=torch.float32)
    check(_lib.load().mvs_relative_projections(ptr(proj_matrices), b, v, ptr(out), stream()), "mvs_relative_projections")
    return out


def relative_projection_pair(src_proj, ref_proj):
    src_proj, ref_proj = _f32(src_proj).contiguous(), _f32(ref_proj).contiguous()
    require_cuda(src_proj, ref_proj)
    b = src_proj.shape[0]
    out = torch.empty(b, 12, device=src_proj.device, dtype=torch.float32)
    check(_lib.load().mvs_relative_projection_pair(ptr(src_proj), ptr(ref_proj), b, ptr(out), stream()),
          "mvs_relative_projection_pair")
    return out


def homo_warp(src_fea, relproj, depth_values, want_mask):
    src_fea, depth_values = _f32(src_fea).contiguous(), _f32(depth_values).contiguous()
    require_cuda(src_fea, relproj, depth_values)
    b, c, h, w = src_fea.shape
    d = depth_values.shape[1]
    is_map = 1 if depth_values.dim() == 4 else 0
    if is_map and tuple(depth_values.shape) != (b, d, h, w):
        raise RuntimeError("depth_values must be [B,D] or [B,D,H,W] matching the feature map")
    warped = torch.empty(b, c, d, h, w, device=src_fea.device, dtype=torch.float32)
    mask = torch.empty(b, d, h, w, device=src_fea.device, dtype=torch.uint8) if want_mask else None
    check(_lib.load().mvs_homo_warp(ptr(src_fea), ptr(relproj), ptr(depth_values), is_map, ptr(warped), ptr(mask),
                                    b, c, d, h, w, stream()), "mvs_homo_warp")
    return warped, (mask.bool() if want_mask else None)


def _feature_strides(features):
    """features [B,V,C,H,W]: each [C,H,W] block must be dense NCHW; batch/view strides are free."""
    b, v, c, h, w = features.shape
    if features.stride(4) != 1 or features.stride(3) != w or features.stride(2) != h * w:
        features = features.contiguous()
    return features, features.stride(0), features.stride(1)


def cost_volume_entropy(features, relproj, depth_values, groups, want_sim):
    features = _f32(features)
    features, bs, vs = _feature_strides(features)
    depth_values = _f32(depth_values).contiguous()
    require_cuda(relproj, depth_values)
    require_cuda_device(features)
    b, v, c, h, w = features.shape
    d = depth_values.shape[1]
    entropy = torch.empty(b, v - 1, h, w, device=features.device, dtype=torch.float32)
    sim = torch.empty(b, d, h, w, device=features.device, dtype=torch.float32) if want_sim else None
    check(_lib.load().mvs_cost_volume_entropy(ptr(features), bs, vs, ptr(relproj), ptr(depth_values), ptr(entropy),
                                              ptr(sim), b, v, c, groups, d, h, w, stream()), "mvs_cost_volume_entropy")
    return entropy, sim


def cost_volume_aggregate(features, relproj, depth_values, vis_weight, groups, round_tf32=False):
    features = _f32(features)
    features, bs, vs = _feature_strides(features)
    depth_values = _f32(depth_values).contiguous()
    require_cuda(relproj, depth_values, vis_weight)
    b, v, c, h, w = features.shape
    d = depth_values.shape[1]
    volume = torch.empty(b, d, h, w, groups, device=features.device, dtype=torch.float32)
    fn = _lib.load().mvs_cost_volume_aggregate_tf32 if round_tf32 else _lib.load().mvs_cost_volume_aggregate
    check(fn(ptr(features), bs, vs, ptr(relproj), ptr(depth_values), ptr(vis_weight), ptr(volume), b, v, c, groups, d, h, w,
             stream()), "mvs_cost_volume_aggregate")
    return volume


def features_to_cl(feature_list, outs=None):
    """NCHW feature tensors ``[..., C, H, W]`` (up to four, e.g. the four stages of one feature set) -> channels-last
    ``[..., H, W, C]`` copies, ONE launch (mvs_features_to_cl).  The cost-volume kernels sample channels-last texels.
    ``outs``: optional preallocated contiguous destinations of the same sizes (e.g. slots of a feature pool)."""
    if not 1 <= len(feature_list) <= 4:
        raise RuntimeError("features_to_cl converts 1..4 tensors per call")
    ins = []
    for t in feature_list:
        t = _f32(t).contiguous()
        require_cuda(t)
        ins.append(t)
    if outs is None:
        outs = [torch.empty(t.shape[:-3] + (t.shape[-2], t.shape[-1], t.shape[-3]), device=t.device, dtype=torch.float32) for t in ins]
    else:
        require_cuda(*outs)
        for t, o in zip(ins, outs):
            if o.numel() != t.numel():
                raise RuntimeError("features_to_cl: destination of %d elements for a source of %d" % (o.numel(), t.numel()))
    n = len(ins)
    arr_p = ctypes.c_void_p * n
    in_p = arr_p(*[t.data_ptr() for t in ins])
    out_p = arr_p(*[t.data_ptr() for t in outs])
    chans = (ctypes.c_int * n)(*[t.shape[-3] for t in ins])
    hw = (ctypes.c_int64 * n)(*[t.shape[-2] * t.shape[-1] for t in ins])
    maps = (ctypes.c_int64 * n)(*[t.numel() // (t.shape[-3] * t.shape[-2] * t.shape[-1]) for t in ins])
    check(_lib.load().mvs_features_to_cl(ctypes.cast(in_p, ctypes.c_void_p), ctypes.cast(out_p, ctypes.c_void_p),
                                         ctypes.cast(chans, ctypes.c_void_p), ctypes.cast(hw, ctypes.c_void_p),
                                         ctypes.cast(maps, ctypes.c_void_p), n, stream()), "mvs_features_to_cl")
    return outs


def cl_supported(chans, ndepth, groups):
    """Shapes the channels-last cost-volume kernels are built for (the reference's four stages)."""
    return groups == 8 and (chans, ndepth) in ((64, 32), (32, 16), (16, 8), (8, 4))


def _cl_views(feat_cl, view_slots):
    """(B, V, H, W, C, nmaps, slot array or None) of a dense [B,V,H,W,C] tensor or a pool [S,H,W,C] + view slots."""
    if view_slots is None:
        b, v, h, w, c = feat_cl.shape
        return b, v, h, w, c, b * v, None
    if feat_cl.dim() != 4:
        raise RuntimeError("with view_slots the features must be a pool [slots,H,W,C]")
    nmaps, h, w, c = feat_cl.shape
    slots = (ctypes.c_int * len(view_slots))(*[int(x) for x in view_slots])
    return 1, len(view_slots), h, w, c, nmaps, slots


def cost_volume_cl_entropy(feat_cl, relproj, depth_values, groups, want_sim, view_slots=None):
    """Pass A over channels-last features [B,V,H,W,C] (or a pool [S,H,W,C] whose maps ``view_slots`` are the views of one
    batch item).  Returns (entropy [B,N,H,W], sim_depth [B,H,W] or None, corr [B,N,D,H,W,G] or None): the per-view
    correlation is stored where C/G >= 2 (one sampling pass); sim_depth is the hypothesis with the largest summed cosine
    similarity (the argmax is taken inside the kernel).  None when the shape is not covered."""
    require_cuda(feat_cl, relproj, depth_values)
    b, v, h, w, c, nmaps, slots = _cl_views(feat_cl, view_slots)
    d = depth_values.shape[1]
    if not cl_supported(c, d, groups):
        return None
    entropy = torch.empty(b, v - 1, h, w, device=feat_cl.device, dtype=torch.float32)
    sim = torch.empty(b, h, w, device=feat_cl.device, dtype=torch.float32) if want_sim else None
    corr = torch.empty(b, v - 1, d, h, w, groups, device=feat_cl.device, dtype=torch.float32) if c // groups >= 2 else None
    rc = _lib.load().mvs_cost_volume_cl_entropy(ptr(feat_cl), nmaps, None if slots is None else ctypes.cast(slots, ctypes.c_void_p),
                                                ptr(relproj), ptr(depth_values), ptr(entropy), ptr(sim), ptr(corr),
                                                b, v, c, groups, d, h, w, stream())
    if rc == 1:
        return None
    check(rc, "mvs_cost_volume_cl_entropy")
    return entropy, sim, corr


def cost_volume_cl_aggregate(feat_cl, relproj, depth_values, vis_weight, groups, round_tf32=False, view_slots=None):
    """Pass B over channels-last features (the stage whose correlation is as large as the warped tensor: C/G = 1)."""
    require_cuda(feat_cl, relproj, depth_values, vis_weight)
    b, v, h, w, c, nmaps, slots = _cl_views(feat_cl, view_slots)
    d = depth_values.shape[1]
    volume = torch.empty(b, d, h, w, groups, device=feat_cl.device, dtype=torch.float32)
    check(_lib.load().mvs_cost_volume_cl_aggregate(ptr(feat_cl), nmaps, None if slots is None else ctypes.cast(slots, ctypes.c_void_p),
                                                   ptr(relproj), ptr(depth_values), ptr(vis_weight), ptr(volume),
                                                   b, v, c, groups, d, h, w, 1 if round_tf32 else 0, stream()),
          "mvs_cost_volume_cl_aggregate")
    return volume


def corr_aggregate(corr, vis_weight, round_tf32=False):
    """corr [B,N,D,H,W,8], vis_weight [B,N,H,W] -> volume channels-last [B,D,H,W,8]."""
    require_cuda(corr, vis_weight)
    b, n, d, h, w, g = corr.shape
    if g != 8:
        raise RuntimeError("corr_aggregate: only G = 8 groups is built")
    volume = torch.empty(b, d, h, w, g, device=corr.device, dtype=torch.float32)
    check(_lib.load().mvs_corr_aggregate(ptr(corr), ptr(vis_weight), ptr(volume), b, n, d, h, w, 1 if round_tf32 else 0,
                                         stream()), "mvs_corr_aggregate")
    return volume


def argmax_gather(score, depth_values):
    require_cuda(score, depth_values)
    b, d, h, w = score.shape
    out = torch.empty(b, h, w, device=score.device, dtype=torch.float32)
    check(_lib.load().mvs_argmax_gather(ptr(score), ptr(depth_values), ptr(out), b, d, h, w, stream()), "mvs_argmax_gather")
    return out


def vis_weight(entropy_maps, params_host):
    """entropy_maps [M,H,W] (device), params_host: float32 numpy array of MVS_VIS_PARAM_FLOATS."""
    require_cuda(entropy_maps)
    m, h, w = entropy_maps.shape
    assert params_host.dtype == np.float32 and params_host.flags["C_CONTIGUOUS"]
    out = torch.empty_like(entropy_maps)
    check(_lib.load().mvs_vis_weight(ptr(entropy_maps), params_host.ctypes.data_as(ctypes.c_void_p), ptr(out), m, h, w,
                                     stream()), "mvs_vis_weight")
    return out


def pack_vis_fused_weights(w, rows):
    """[Cout, 16, 3, 3] (BN folded) -> [kw][4 quads][rows][4] TF32: per (kw, input-channel quad) the B rows are [kh][cout]
    (3 * Cout of them, zero-padded to ``rows``): the three kernel rows are one MMA operand (mvs_vis_fused)."""
    cout, cin = w.shape[:2]
    t = w.permute(3, 1, 2, 0).reshape(3, cin // 4, 4, 3, cout).permute(0, 1, 3, 4, 2).reshape(3, cin // 4, 3 * cout, 4)
    if rows > 3 * cout:
        t = torch.nn.functional.pad(t, (0, 0, 0, rows - 3 * cout))
    return round_tf32(t.contiguous())


def vis_fused(entropy_maps, params_host, w2_packed, w3_packed):
    """[M,H,W] entropy -> [M,H,W] visibility weight, the whole net in one persistent kernel (mvs_vis_fused)."""
    require_cuda(entropy_maps, w2_packed, w3_packed)
    m, h, w = entropy_maps.shape
    assert params_host.dtype == np.float32 and params_host.flags["C_CONTIGUOUS"] and params_host.size == 193
    out = torch.empty_like(entropy_maps)
    check(_lib.load().mvs_vis_fused(ptr(entropy_maps), params_host.ctypes.data_as(ctypes.c_void_p), ptr(w2_packed), ptr(w3_packed),
                                    ptr(out), m, h, w, stream()), "mvs_vis_fused")
    return out


def conv3d_cl(x, w_packed, shift, skip, stride, relu=True):
    """x [B,D,H,W,Cin] -> [B,Do,Ho,Wo,Cout]; w_packed [kd,3,3,Cin,Cout]."""
    require_cuda(x, w_packed, shift, skip)
    b, d, h, w, cin = x.shape
    kd, cout = w_packed.shape[0], w_packed.shape[4]
    sd, sh, sw = stride
    pd = kd // 2
    do, ho, wo = (d + 2 * pd - kd) // sd + 1, (h - 1) // sh + 1, (w - 1) // sw + 1
    y = torch.empty(b, do, ho, wo, cout, device=x.device, dtype=torch.float32)
    if skip is not None and tuple(skip.shape) != tuple(y.shape):
        raise RuntimeError("The size of tensor a %s must match the size of tensor b %s (skip connection)"
                           % (tuple(skip.shape), tuple(y.shape)))
    check(_lib.load().mvs_conv3d_cl(ptr(x), ptr(w_packed), ptr(shift), ptr(skip), ptr(y), b, d, h, w, cin, cout, kd,
                                    sd, sh, sw, 1 if relu else 0, stream()), "mvs_conv3d_cl")
    return y


def deconv3d_cl(x, w_packed, shift, skip, sd, relu=True):
    """ConvTranspose3d (kd,3,3), stride (sd,2,2), pad k//2, output_padding stride-1."""
    require_cuda(x, w_packed, shift, skip)
    b, d, h, w, cin = x.shape
    kd, cout = w_packed.shape[0], w_packed.shape[4]
    y = torch.empty(b, d * sd, h * 2, w * 2, cout, device=x.device, dtype=torch.float32)
    if skip is not None and tuple(skip.shape) != tuple(y.shape):
        raise RuntimeError("The size of tensor a %s must match the size of tensor b %s (skip connection)"
                           % (tuple(skip.shape), tuple(y.shape)))
    check(_lib.load().mvs_deconv3d_cl(ptr(x), ptr(w_packed), ptr(shift), ptr(skip), ptr(y), b, d, h, w, cin, cout, kd,
                                      sd, 1 if relu else 0, stream()), "mvs_deconv3d_cl")
    return y


def ncdhw_to_cl(x, round_tf32=False):
    x = _f32(x).contiguous()
    require_cuda(x)
    b, c, d, h, w = x.shape
    y = torch.empty(b, d, h, w, c, device=x.device, dtype=torch.float32)
    fn = _lib.load().mvs_ncdhw_to_cl_tf32 if round_tf32 else _lib.load().mvs_ncdhw_to_cl
    check(fn(ptr(x), ptr(y), b, c, d, h, w, stream()), "mvs_ncdhw_to_cl")
    return y


def cl_to_ncdhw(x):
    require_cuda(x)
    b, d, h, w, c = x.shape
    y = torch.empty(b, c, d, h, w, device=x.device, dtype=torch.float32)
    check(_lib.load().mvs_cl_to_ncdhw(ptr(x), ptr(y), b, c, d, h, w, stream()), "mvs_cl_to_ncdhw")
    return y


def prob_conv_cl(x, w_host, bias_host, ksize):
    require_cuda(x)
    b, d, h, w, cin = x.shape
    pre = torch.empty(b, d, h, w, device=x.device, dtype=torch.float32)
    bias_p = bias_host.ctypes.data_as(ctypes.c_void_p) if bias_host is not None else None
    check(_lib.load().mvs_prob_conv_cl(ptr(x), w_host.ctypes.data_as(ctypes.c_void_p), bias_p, ptr(pre), b, d, h, w, cin,
                                       ksize, stream()), "mvs_prob_conv_cl")
    return pre


def regression_head(pre, depth_values, tmp, training, want_prob=True):
    pre, depth_values = _f32(pre).contiguous(), _f32(depth_values).contiguous()
    require_cuda(pre, depth_values)
    b, d, h, w = pre.shape
    if tuple(depth_values.shape) != (b, d, h, w):
        raise RuntimeError("depth_values %s does not match prob volume %s" % (tuple(depth_values.shape), tuple(pre.shape)))
    prob = torch.empty_like(pre) if want_prob else None
    depth = torch.empty(b, h, w, device=pre.device, dtype=torch.float32)
    conf = torch.empty(b, h, w, device=pre.device, dtype=torch.float32)
    check(_lib.load().mvs_regression_head(ptr(pre), ptr(depth_values), float(tmp), 1 if training else 0, ptr(prob),
                                          ptr(depth), ptr(conf), b, d, h, w, stream()), "mvs_regression_head")
    return prob, depth, conf


def depth_regression(p, depth_values):
    p, depth_values = _f32(p).contiguous(), _f32(depth_values).contiguous()
    require_cuda(p, depth_values)
    b, d, h, w = p.shape
    # the reference broadcasts: [D] (conf_regression's arange), [B,D], [B,D,1,1] or a full [B,D,H,W] map (module.py:597-603)
    if depth_values.dim() == 1:
        depth_values = depth_values.unsqueeze(0).expand(b, -1).contiguous()
    elif depth_values.dim() == 4 and tuple(depth_values.shape[2:]) == (1, 1):
        depth_values = depth_values.reshape(depth_values.shape[0], -1)
    if depth_values.dim() == 2 and depth_values.shape[0] == 1 and b > 1:
        depth_values = depth_values.expand(b, -1).contiguous()
    if tuple(depth_values.shape) not in ((b, d), (b, d, h, w)):
        raise RuntimeError("depth_regression: depth_values %s do not broadcast against p %s"
                           % (tuple(depth_values.shape), tuple(p.shape)))
    is_map = 1 if depth_values.dim() == 4 else 0
    out = torch.empty(b, h, w, device=p.device, dtype=torch.float32)
    check(_lib.load().mvs_depth_regression(ptr(p), ptr(depth_values), is_map, ptr(out), b, d, h, w, stream()),
          "mvs_depth_regression")
    return out


def conf_regression(p, n):
    p = _f32(p).contiguous()
    require_cuda(p)
    b, d, h, w = p.shape
    out = torch.empty(b, h, w, device=p.device, dtype=torch.float32)
    check(_lib.load().mvs_conf_regression(ptr(p), int(n), ptr(out), b, d, h, w, stream()), "mvs_conf_regression")
    return out


def init_range(cur_depth, ndepths, h, w, inverse):
    cur_depth = _f32(cur_depth).contiguous()
    require_cuda(cur_depth)
    b, nd = cur_depth.shape
    out = torch.empty(b, ndepths, h, w, device=cur_depth.device, dtype=torch.float32)
    fn = _lib.load().mvs_init_inverse_range if inverse else _lib.load().mvs_init_range
    check(fn(ptr(cur_depth), nd, ptr(out), b, ndepths, h, w, stream()), "mvs_init_range")
    return out


def schedule_inverse_range(depth, depth_hypo, ndepths, split_itv, h, w):
    depth, depth_hypo = _f32(depth).contiguous(), _f32(depth_hypo).contiguous()
    require_cuda(depth, depth_hypo)
    b, dprev, h2, w2 = depth_hypo.shape
    if (h2, w2) != (h // 2, w // 2) or tuple(depth.shape) != (b, h2, w2):
        raise RuntimeError("schedule_inverse_range: previous-stage maps must be [B,%d,%d]" % (h // 2, w // 2))
    out = torch.empty(b, ndepths, h, w, device=depth.device, dtype=torch.float32)
    check(_lib.load().mvs_schedule_inverse_range(ptr(depth), ptr(depth_hypo), dprev, float(split_itv), ptr(out), b,
                                                 ndepths, h, w, stream()), "mvs_schedule_inverse_range")
    return out


def schedule_range(cur_depth, ndepth, depth_interval_pixel, h, w):
    cur_depth, itv = _f32(cur_depth).contiguous(), _f32(depth_interval_pixel).contiguous()
    require_cuda(cur_depth, itv)
    b = cur_depth.shape[0]
    if tuple(cur_depth.shape) != (b, h // 2, w // 2):
        raise RuntimeError("schedule_range: the previous-stage depth must be [B,%d,%d] (got %s)"
                           % (h // 2, w // 2, tuple(cur_depth.shape)))
    out = torch.empty(b, ndepth, h, w, device=cur_depth.device, dtype=torch.float32)
    check(_lib.load().mvs_schedule_range(ptr(cur_depth), ptr(itv), ptr(out), b, ndepth, h, w, stream()), "mvs_schedule_range")
    return out


def confidence_upsample_accumulate(conf, acc, scale=1.0):
    """acc += scale * nearest_upsample(conf); returns the upsampled stage confidence."""
    require_cuda(conf, acc)
    b, h, w = conf.shape
    up = torch.empty_like(acc)
    check(_lib.load().mvs_confidence_upsample_accumulate(ptr(conf), h, w, ptr(up), ptr(acc), b, acc.shape[1], acc.shape[2],
                                                         float(scale), stream()), "mvs_confidence_upsample_accumulate")
    return up


def confidence_accumulate(conf, acc, scale=1.0):
    require_cuda(conf, acc)
    b, h, w = conf.shape
    check(_lib.load().mvs_confidence_accumulate(ptr(conf), h, w, ptr(acc), b, acc.shape[1], acc.shape[2], float(scale),
                                                stream()), "mvs_confidence_accumulate")
    return acc


# ------------------------------------------------------------------------------------------------
# tensor-core (tcgen05) convolution path
# ------------------------------------------------------------------------------------------------
def round_tf32(t):
    """Round-to-nearest (ties away) to TF32, the same as PTX cvt.rna.tf32.f32."""
    bits = t.contiguous().view(torch.int32)
    return ((bits + 0x1000) & -8192).view(torch.float32)


def tc_channel_slice(cin):
    return 32 if cin >= 32 else cin


def tc_n_tile(cout, x3=False, transposed=False):
    if cout <= 16:
        return 16
    if cout == 32 or x3 or transposed:     # 3xTF32 doubles the staged operands: keep two stages in smem
        return 32
    return 64


def pack_tc_weights(w_packed, x3):
    """w_packed [kd,3,3,Cin,Cout] (BN folded) -> operand-order arrays for mvs_conv3d_tc:
    [Cout_tiles][kd][kh][Cin/CS][kw][CS/4][n_tile][4]; returns (w_hi, w_lo or None, n_tile)."""
    kd, _, _, cin, cout = w_packed.shape
    cs, nt = tc_channel_slice(cin), tc_n_tile(cout, x3)
    ntiles = (cout + nt - 1) // nt
    w = w_packed
    if ntiles * nt != cout:
        w = torch.nn.functional.pad(w, (0, ntiles * nt - cout))
    w = w.reshape(kd, 3, 3, cin // cs, cs // 4, 4, ntiles, nt).permute(6, 0, 1, 3, 2, 4, 7, 5).contiguous()
    hi = round_tf32(w)
    lo = round_tf32(w - hi) if x3 else None
    return hi, lo, nt


def pack_tc_deconv_weights(w_packed, x3):
    """w_packed [kd,3,3,Cin,Cout] (from torch's ConvTranspose3d [Cin,Cout,kd,kh,kw], BN folded) ->
    [Cout_tiles][kd][2 dy][Cin/CS][6 taps][CS/4][n_tile][4] for mvs_deconv3d_tc."""
    kd, _, _, cin, cout = w_packed.shape
    cs, nt = tc_channel_slice(cin), tc_n_tile(cout, x3, transposed=True)
    ntiles = (cout + nt - 1) // nt
    w = w_packed
    if ntiles * nt != cout:
        w = torch.nn.functional.pad(w, (0, ntiles * nt - cout))
    w = w.reshape(kd, 3, 3, cin // cs, cs // 4, 4, ntiles, nt)               # [kz, kh, kw, ch, q, e, tile, n]
    out = torch.zeros(ntiles, kd, 2, cin // cs, 6, cs // 4, nt, 4, device=w.device, dtype=w.dtype)
    taps = {0: [(1, 0), (1, 1), (1, 2), (2, 0), (2, 1), (2, 2)], 1: [(0, 0), (0, 1), (0, 2)]}
    for dy, lst in taps.items():
        for t, (kh, kw) in enumerate(lst):
            out[:, :, dy, :, t] = w[:, kh, kw].permute(4, 0, 1, 2, 5, 3)      # [tile, kz, ch, q, n, e]
    hi = round_tf32(out.contiguous())
    lo = round_tf32(out - hi) if x3 else None
    return hi, lo, nt


def deconv3d_tc(x, w_hi, w_lo, n_tile, cout, kd, shift, skip, sd, relu=True):
    require_cuda(x, w_hi, w_lo, shift, skip)
    b, d, h, w, cin = x.shape
    y = torch.empty(b, d * sd, h * 2, w * 2, cout, device=x.device, dtype=torch.float32)
    if skip is not None and tuple(skip.shape) != tuple(y.shape):
        raise RuntimeError("The size of tensor a %s must match the size of tensor b %s (skip connection)"
                           % (tuple(skip.shape), tuple(y.shape)))
    check(_lib.load().mvs_deconv3d_tc(ptr(x), ptr(w_hi), ptr(w_lo), ptr(shift), ptr(skip), ptr(y), b, d, h, w, cin, cout,
                                      n_tile, kd, sd, 1 if relu else 0, stream()), "mvs_deconv3d_tc")
    return y


def conv3d_tc(x, w_hi, w_lo, n_tile, cout, kd, shift, skip, stride, relu=True):
    require_cuda(x, w_hi, w_lo, shift, skip)
    b, d, h, w, cin = x.shape
    sd, sh, sw = stride
    if sh != sw:
        raise RuntimeError("conv3d_tc: H and W strides must match")
    pd = kd // 2
    do, ho, wo = (d + 2 * pd - kd) // sd + 1, (h - 1) // sh + 1, (w - 1) // sw + 1
    y = torch.empty(b, do, ho, wo, cout, device=x.device, dtype=torch.float32)
    if skip is not None and tuple(skip.shape) != tuple(y.shape):
        raise RuntimeError("The size of tensor a %s must match the size of tensor b %s (skip connection)"
                           % (tuple(skip.shape), tuple(y.shape)))
    check(_lib.load().mvs_conv3d_tc(ptr(x), ptr(w_hi), ptr(w_lo), ptr(shift), ptr(skip), ptr(y), b, d, h, w, cin, cout,
                                    n_tile, kd, sd, sh, 1 if relu else 0, stream()), "mvs_conv3d_tc")
    return y


def tc_probe(a_img, b_img, a_lbo, a_sbo, b_lbo, b_sbo, n, nk, a_kstep, b_kstep):
    require_cuda(a_img, b_img)
    out = torch.empty(128, n, device=a_img.device, dtype=torch.float32)
    check(_lib.load().mvs_tc_probe(ptr(a_img), a_img.numel() * 4, ptr(b_img), b_img.numel() * 4, a_lbo, a_sbo, b_lbo, b_sbo,
                                   n, nk, a_kstep, b_kstep, ptr(out), stream()), "mvs_tc_probe")
    return out


# ------------------------------------------------------------------------------------------------
# round-2 persistent TMA-fed tcgen05 convolutions (conv3d_tma.cu)
# ------------------------------------------------------------------------------------------------
TMA_S1, TMA_S2, TMA_DECONV = 0, 1, 2
_TMA_BUILT = {(0, 16, 16), (0, 32, 32), (0, 64, 16), (1, 8, 16), (1, 16, 32), (1, 32, 16), (2, 64, 16), (2, 32, 16), (2, 16, 16)}


def tma_n_tile(cin, cout, mode=TMA_S1):
    """N tile such that the layer's whole weight tile (27 x Cin x n_tile floats) and the input stages fit shared memory."""
    if mode == TMA_DECONV or cout <= 16 or cin >= 64 or (mode == TMA_S2 and cin >= 32):
        return 16
    return 32


def tma_supported(cin, cout, d, kd, stride2=False, transposed=False):
    """Shapes mvs_conv3d_tma is built for (depth-unstrided layers)."""
    if kd not in (1, 3) or cout % 4:
        return False
    mode = TMA_DECONV if transposed else (TMA_S2 if stride2 else TMA_S1)
    nt = tma_n_tile(cin, cout, mode)
    return (mode, cin, nt) in _TMA_BUILT and (4 if transposed else 1) * d * nt <= 512


def pack_tma_weights(w_packed, mode=TMA_S1):
    """[kd,3,3,Cin,Cout] -> [Cout_tiles][kh][kw][Cin/4][kd][n_tile][4], TF32-rounded: per (tap, channel quad) the B rows are
    [kz][n], so the depth taps of a slab are one operand of N = kd * n_tile rows (mvs_conv3d_tma).  For the transposed mode
    pass torch's ConvTranspose3d weight permuted to [kd,kh,kw,Cin,Cout]."""
    kd, _, _, cin, cout = w_packed.shape
    nt = tma_n_tile(cin, cout, mode)
    ntiles = (cout + nt - 1) // nt
    w = w_packed
    if ntiles * nt != cout:
        w = torch.nn.functional.pad(w, (0, ntiles * nt - cout))
    #            [kz, kh, kw, q, e, tile, n]  ->  [tile, kh, kw, q, kz, n, e]
    w = w.reshape(kd, 3, 3, cin // 4, 4, ntiles, nt).permute(5, 1, 2, 3, 0, 6, 4).contiguous()
    return round_tf32(w), nt


def conv3d_tma(x, w_tma, n_tile, cout, kd, shift, skip, relu=True, mode=TMA_S1):
    require_cuda(x, w_tma, shift, skip)
    b, d, h, w, cin = x.shape
    ho, wo = ((h + 1) // 2, (w + 1) // 2) if mode == TMA_S2 else ((2 * h, 2 * w) if mode == TMA_DECONV else (h, w))
    y = torch.empty(b, d, ho, wo, cout, device=x.device, dtype=torch.float32)
    if skip is not None and tuple(skip.shape) != tuple(y.shape):
        raise RuntimeError("The size of tensor a %s must match the size of tensor b %s (skip connection)"
                           % (tuple(skip.shape), tuple(y.shape)))
    check(_lib.load().mvs_conv3d_tma(ptr(x), ptr(w_tma), ptr(shift), ptr(skip), ptr(y), b, d, h, w, cin, cout, n_tile, kd, mode,
                                     1 if relu else 0, stream()), "mvs_conv3d_tma")
    return y


def conv3d_tma_prob(x, w_tma, kd, shift, skip, prob_w_host, prob_bias, relu=True):
    """Transposed conv 16 -> 8 (mode TMA_DECONV, + shift, ReLU, + skip) with the 1x1x1 ``prob`` conv 8 -> 1 in its epilogue
    (mvs_conv3d_tma_prob): x [B,D,H,W,16] -> prob_volume_pre [B,D,2H,2W]; prob_w_host: float32 numpy [8]."""
    require_cuda(x, w_tma, shift, skip)
    b, d, h, w, cin = x.shape
    if skip is not None and tuple(skip.shape) != (b, d, 2 * h, 2 * w, 8):
        raise RuntimeError("The size of tensor a %s must match the size of tensor b %s (skip connection)"
                           % (tuple(skip.shape), (b, d, 2 * h, 2 * w, 8)))
    if prob_w_host.size != 8:
        raise RuntimeError("prob weights must hold 8 values, got %d" % prob_w_host.size)
    pre = torch.empty(b, d, 2 * h, 2 * w, device=x.device, dtype=torch.float32)
    check(_lib.load().mvs_conv3d_tma_prob(ptr(x), ptr(w_tma), ptr(shift), ptr(skip), prob_w_host.ctypes.data_as(ctypes.c_void_p),
                                          float(prob_bias), ptr(pre), b, d, h, w, cin, kd, 1 if relu else 0, stream()),
          "mvs_conv3d_tma_prob")
    return pre


def tc_probe_ts(a_img, b_img, a_lbo, a_sbo, b_lbo, b_sbo, n, nk, a_kstep, b_kstep, a_shift_bytes=0):
    require_cuda(a_img, b_img)
    out = torch.empty(128, n, device=a_img.device, dtype=torch.float32)
    check(_lib.load().mvs_tc_probe_ts(ptr(a_img), a_img.numel() * 4, ptr(b_img), b_img.numel() * 4, a_lbo, a_sbo, b_lbo, b_sbo,
                                      n, nk, a_kstep, b_kstep, a_shift_bytes, ptr(out), stream()), "mvs_tc_probe_ts")
    return out
