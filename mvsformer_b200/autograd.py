"""Training path of the drop-in modules: train-mode forward (batch-statistics BatchNorm) and the
backward of StageNet, as ``torch.autograd.Function`` wrappers over the C ABI (SURVEY.md §8b
"Autograd"; BASELINE cfg 5).

What the reference differentiates (models/mvsformer_model.py:61-125, training branch):

* features (reference and source views) through the bilinear gather and the group-wise
  correlation — the sampling grid is built under ``no_grad`` (models/warping.py:79), so cameras
  and depth hypotheses receive nothing;
* the visibility net's parameters through ``vis_weight`` — its entropy input is detached (:88);
* every ``cost_reg`` parameter; the loss enters through ``prob_volume_pre`` (models/losses.py:311-341),
  ``prob_volume`` stays differentiable, ``depth`` (argmax gather) and the confidence are not.

All arithmetic runs in libmvs_b200.so (``csrc/train.cu`` plus the FP32 convolution kernels of
``csrc/conv3d.cu``, which also serve as data-gradient kernels with re-packed weights).  torch is
used for memory, tiny weight re-layouts (permute / flip of <= 110k-element tensors), the
SyncBatchNorm all-reduce and autograd bookkeeping.  This first version is FP32 on CUDA cores,
whatever ``MVS_CONV_PRECISION`` says; there is no fallback: CPU tensors raise.
"""
import torch
import torch.distributed as dist
import torch.nn as nn

from . import _lib, config, engine


_PROFILE = None          # {entry point: [(start event, end event), ...]} while profile_kernels() is active


class profile_kernels:
    """``with profile_kernels() as prof: step()`` then ``prof.summary()`` -> {entry point: (ms, launches)} from
    CUDA events recorded around every training-path launch on the current stream (bench / scripts only)."""

    def __enter__(self):
        global _PROFILE
        _PROFILE = self.events = {}
        return self

    def __exit__(self, *exc):
        global _PROFILE
        _PROFILE = None

    def summary(self):
        torch.cuda.synchronize()
        return {k: (sum(a.elapsed_time(b) for a, b in v), len(v)) for k, v in self.events.items()}


class _timed:
    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if _PROFILE is not None:
            self.start = torch.cuda.Event(enable_timing=True)
            self.start.record()

    def __exit__(self, *exc):
        if _PROFILE is not None:
            end = torch.cuda.Event(enable_timing=True)
            end.record()
            _PROFILE.setdefault(self.name, []).append((self.start, end))


def _call(name, *args):
    with _timed(name):
        _lib.check(getattr(_lib.load(), name)(*args, _lib.stream()), name)


def _p(t):
    return _lib.ptr(t)


# ------------------------------------------------------------------------------------------------
# tensor-level wrappers (no autograd)
# ------------------------------------------------------------------------------------------------


def group_corr_fwd(features, relproj, depth, groups):
    """features [B,V,C,H,W], relproj [B,N,12], depth [B,D,H,W] -> corr [B,N,D,H,W,G]."""
    features, bs, vs = engine._feature_strides(features)
    _lib.require_cuda(features, relproj, depth)
    b, v, c, h, w = features.shape
    d = depth.shape[1]
    corr = torch.empty(b, v - 1, d, h, w, groups, device=features.device, dtype=torch.float32)
    _call("mvs_group_corr_fwd", _p(features), bs, vs, _p(relproj), _p(depth), _p(corr), b, v, c, groups, d, h, w)
    return corr


def group_corr_bwd(features, relproj, depth, gcorr, groups):
    features, bs, vs = engine._feature_strides(features)
    gcorr = gcorr.contiguous()
    _lib.require_cuda(features, relproj, depth, gcorr)
    b, v, c, h, w = features.shape
    d = depth.shape[1]
    gfeat = torch.zeros(b, v, c, h, w, device=features.device, dtype=torch.float32)
    _call("mvs_group_corr_bwd", _p(features), bs, vs, _p(relproj), _p(depth), _p(gcorr), _p(gfeat), b, v, c, groups, d, h, w)
    return gfeat


def corr_entropy(corr):
    """corr [B,N,D,H,W,G] -> entropy [B,N,H,W] (mvsformer_model.py:87-90)."""
    _lib.require_cuda(corr)
    b, n, d, h, w, g = corr.shape
    out = torch.empty(b, n, h, w, device=corr.device, dtype=torch.float32)
    _call("mvs_corr_entropy", _p(corr), _p(out), b * n, g, d, h, w)
    return out


def aggregate_fwd(corr, weight):
    _lib.require_cuda(corr, weight)
    b, n, d, h, w, g = corr.shape
    vol = torch.empty(b, d, h, w, g, device=corr.device, dtype=torch.float32)
    _call("mvs_aggregate_fwd", _p(corr), _p(weight), _p(vol), b, n, g, d, h, w)
    return vol


def aggregate_bwd(gvol, corr, weight):
    gvol = gvol.contiguous()
    _lib.require_cuda(gvol, corr, weight)
    b, n, d, h, w, g = corr.shape
    gcorr = torch.empty_like(corr)
    gweight = torch.empty_like(weight)
    _call("mvs_aggregate_bwd", _p(gvol), _p(corr), _p(weight), _p(gcorr), _p(gweight), b, n, g, d, h, w)
    return gcorr, gweight


BN_REPLICAS = 32          # MVS_BN_REPLICAS of include/mvs_b200.h


def _reduction_buffer(channels, device):
    """Zeroed accumulators of the BatchNorm reductions: BN_REPLICAS copies of 2C doubles (see mvs_bn_stats)."""
    return torch.zeros(BN_REPLICAS * 2 * channels, device=device, dtype=torch.float64)


def channel_sums(x, channels):
    """x [M,C] channels-last (any leading shape) -> float64 [2C]: per-channel sum and sum of squares."""
    _lib.require_cuda(x)
    sums = _reduction_buffer(channels, x.device)
    _call("mvs_bn_stats", _p(x), _p(sums), x.numel() // channels, channels)
    _call("mvs_bn_collapse", _p(sums), channels)
    return sums[:2 * channels]


def thin_conv(x, w_taps, bias, kd, khw, act):
    """x [B,D,H,W,Cin], w_taps [kd*k*k,Cin,Cout] (device) -> [B,D,H,W,Cout]."""
    _lib.require_cuda(x, w_taps, bias)
    b, d, h, w, cin = x.shape
    cout = w_taps.shape[2]
    y = torch.empty(b, d, h, w, cout, device=x.device, dtype=torch.float32)
    _call("mvs_thin_conv_cl", _p(x), _p(w_taps), _p(bias), _p(y), b, d, h, w, cin, cout, kd, khw, act)
    return y


def conv_wgrad(small, big, kd, khw, sd, shw, small_is_cout):
    """Packed weight gradient [kd,k,k,Cin,Cout] (see mvs_conv_wgrad_cl)."""
    small, big = small.contiguous(), big.contiguous()
    _lib.require_cuda(small, big)
    b, ds, hs, ws, cs = small.shape
    _, db, hb, wb, cb = big.shape
    cin, cout = (cb, cs) if small_is_cout else (cs, cb)
    dw = torch.zeros(kd, khw, khw, cin, cout, device=small.device, dtype=torch.float32)
    _call("mvs_conv_wgrad_cl", _p(small), _p(big), _p(dw), b, ds, hs, ws, db, hb, wb, cs, cb, kd, khw, sd, shw,
          1 if small_is_cout else 0)
    return dw


# ------------------------------------------------------------------------------------------------
# cost volume
# ------------------------------------------------------------------------------------------------


class _GroupCorrelation(torch.autograd.Function):
    @staticmethod
    def forward(ctx, features, relproj, depth, groups):
        ctx.save_for_backward(features, relproj, depth)
        ctx.groups = groups
        return group_corr_fwd(features, relproj, depth, groups)

    @staticmethod
    def backward(ctx, gcorr):
        features, relproj, depth = ctx.saved_tensors
        gfeat = group_corr_bwd(features, relproj, depth, gcorr, ctx.groups) if ctx.needs_input_grad[0] else None
        return gfeat, None, None, None


def group_correlation(features, relproj, depth_values, groups):
    """Differentiable (w.r.t. features) per-view group correlation [B,N,D,H,W,G]."""
    return _GroupCorrelation.apply(engine._f32(features), relproj, depth_values, groups)


class _Aggregate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, corr, weight):
        ctx.save_for_backward(corr, weight)
        return aggregate_fwd(corr, weight)

    @staticmethod
    def backward(ctx, gvol):
        corr, weight = ctx.saved_tensors
        gcorr, gweight = aggregate_bwd(gvol, corr, weight)
        return gcorr, gweight


def aggregate(corr, weight):
    """corr [B,N,D,H,W,G], weight [B,N,H,W] -> volume_mean channels-last [B,D,H,W,G]."""
    return _Aggregate.apply(corr.contiguous(), weight.contiguous())


class _HomoWarp(torch.autograd.Function):
    """homo_warping_3D[_with_mask] (models/warping.py:69-109): differentiable w.r.t. src_fea, as the
    reference's F.grid_sample is; the grid (cameras, depth) is built under no_grad there (:79)."""

    @staticmethod
    def forward(ctx, src_fea, relproj, depth_values, want_mask):
        warped, mask = engine.homo_warp(src_fea, relproj, depth_values, want_mask)
        ctx.save_for_backward(relproj, depth_values)
        ctx.src_shape = tuple(src_fea.shape)
        if mask is not None:
            ctx.mark_non_differentiable(mask)
            return warped, mask
        return warped, None

    @staticmethod
    def backward(ctx, gwarped, _gmask):
        relproj, depth_values = ctx.saved_tensors
        if not ctx.needs_input_grad[0]:
            return None, None, None, None
        gwarped = gwarped.contiguous()
        depth_values = depth_values.float().contiguous()
        _lib.require_cuda(gwarped, relproj, depth_values)
        b, c, h, w = ctx.src_shape
        d = depth_values.shape[1]
        gsrc = torch.zeros(b, c, h, w, device=gwarped.device, dtype=torch.float32)
        _call("mvs_homo_warp_bwd", _p(gwarped), _p(relproj), _p(depth_values), 1 if depth_values.dim() == 4 else 0,
              _p(gsrc), b, c, d, h, w)
        return gsrc, None, None, None


def homo_warp(src_fea, relproj, depth_values, want_mask):
    return _HomoWarp.apply(engine._f32(src_fea), relproj, depth_values, want_mask)


WARP_GRAD_REPLICAS = 32          # MVS_WARP_GRAD_REPLICAS of include/mvs_b200.h


class _DiffHomoWarp(torch.autograd.Function):
    """diff_homo_warping_3D_with_mask (models/warping.py:112-152): the sampling grid is part of the graph, so the
    relative projection [R|t] (-> cameras, through torch's 4x4 algebra) and the depth hypotheses receive gradients too."""

    @staticmethod
    def forward(ctx, src_fea, relproj, depth_values):
        warped, mask = engine.homo_warp(src_fea, relproj, depth_values, True)
        ctx.save_for_backward(src_fea, relproj, depth_values)
        ctx.mark_non_differentiable(mask)
        return warped, mask

    @staticmethod
    def backward(ctx, gwarped, _gmask):
        src_fea, relproj, depth_values = ctx.saved_tensors
        gwarped = gwarped.contiguous()
        src_fea, dv = src_fea.contiguous(), depth_values.float().contiguous()
        _lib.require_cuda(gwarped, src_fea, relproj, dv)
        b, c, h, w = src_fea.shape
        d = dv.shape[1]
        is_map = 1 if dv.dim() == 4 else 0
        gsrc = grel = gdep = None
        if ctx.needs_input_grad[0]:
            gsrc = torch.zeros(b, c, h, w, device=gwarped.device, dtype=torch.float32)
            _call("mvs_homo_warp_bwd", _p(gwarped), _p(relproj), _p(dv), is_map, _p(gsrc), b, c, d, h, w)
        if ctx.needs_input_grad[1] or ctx.needs_input_grad[2]:
            gdep = torch.zeros_like(dv)
            grep = torch.zeros(b, WARP_GRAD_REPLICAS, 12, device=gwarped.device, dtype=torch.float32)
            _call("mvs_homo_warp_bwd_grid", _p(gwarped), _p(src_fea), _p(relproj), _p(dv), is_map, _p(gdep), _p(grep),
                  b, c, d, h, w)
            grel = grep.sum(dim=1)
        return gsrc, grel if ctx.needs_input_grad[1] else None, gdep if ctx.needs_input_grad[2] else None


def diff_homo_warp(src_fea, src_proj, ref_proj, depth_values):
    """The projection algebra stays torch's (:122-124, differentiable 4x4 matmul / inverse); warp and gradients are ours."""
    proj = torch.matmul(src_proj.float(), torch.inverse(ref_proj.float()))
    relproj = torch.cat([proj[:, :3, :3], proj[:, :3, 3:4]], dim=2).reshape(-1, 12).contiguous()
    return _DiffHomoWarp.apply(engine._f32(src_fea).contiguous(), relproj, depth_values)


# ------------------------------------------------------------------------------------------------
# conv / transposed conv  ->  BatchNorm (batch statistics)  ->  ReLU  (+ skip)
# ------------------------------------------------------------------------------------------------


def _pack(weight, transposed):
    """torch layout -> packed [kd,kh,kw,Cin,Cout]."""
    return (weight.permute(2, 3, 4, 0, 1) if transposed else weight.permute(2, 3, 4, 1, 0)).contiguous()


def _unpack(dwp, transposed):
    """packed [kd,kh,kw,Cin,Cout] -> torch layout, with canonical strides (DDP's bucket views expect them)."""
    src = dwp.permute(3, 4, 0, 1, 2) if transposed else dwp.permute(4, 3, 0, 1, 2)
    return torch.empty(src.shape, device=dwp.device, dtype=dwp.dtype).copy_(src)


def _fat(cin, cout):
    """True when mvs_conv3d_cl / mvs_deconv3d_cl accept this channel pair."""
    return cin % 4 == 0 and cout % 8 == 0


_TC_CONV_TILES = {(8, 16), (16, 16), (16, 32), (32, 16), (32, 32), (32, 64)}     # (channel slice, N tile) in conv3d_tc.cu
_TC_DECONV_TILES = {(16, 16), (32, 16), (32, 32)}


def _tensor_core_conv(x, wp, transposed, stride):
    """Opt-in (config.train_conv): the same convolution through the tcgen05 kernels of the inference path, or None
    when the layer shape has no instantiation."""
    mode = config.train_conv()
    kd, kh, _, cin, cout = wp.shape
    if mode == "fp32" or kh != 3 or not (cin in (8, 16) or cin % 32 == 0) or cout % 8 or cout > 64:
        return None
    x3 = mode == "tf32x3"
    key = (engine.tc_channel_slice(cin), engine.tc_n_tile(cout, x3, transposed))
    if transposed:
        if key not in _TC_DECONV_TILES:
            return None
        hi, lo, nt = engine.pack_tc_deconv_weights(wp, x3)
        with _timed("mvs_deconv3d_tc"):
            return engine.deconv3d_tc(x, hi, lo, nt, cout, kd, None, None, stride[0], relu=False)
    if key not in _TC_CONV_TILES or stride[1] != stride[2]:
        return None
    hi, lo, nt = engine.pack_tc_weights(wp, x3)
    with _timed("mvs_conv3d_tc"):
        return engine.conv3d_tc(x, hi, lo, nt, cout, kd, None, None, stride, relu=False)


def _raw_conv(x, wp, transposed, stride):
    """Convolution without bias / BN / activation on channels-last x; wp packed."""
    kd, kh, _, cin, cout = wp.shape
    y = _tensor_core_conv(x, wp, transposed, stride)
    if y is not None:
        return y
    if transposed:
        if not _fat(cin, cout):
            raise NotImplementedError("transposed conv %d->%d channels is not built" % (cin, cout))
        with _timed("mvs_deconv3d_cl"):
            return engine.deconv3d_cl(x, wp, None, None, stride[0], relu=False)
    if kh == 3 and _fat(cin, cout):
        with _timed("mvs_conv3d_cl"):
            return engine.conv3d_cl(x, wp, None, None, stride, relu=False)
    if tuple(stride) != (1, 1, 1):
        raise NotImplementedError("strided conv %d->%d channels (kernel %d) is not built" % (cin, cout, kh))
    return thin_conv(x, wp.view(-1, cin, cout), None, kd, kh, 0)


def _conv_dgrad(g, wp, transposed, stride, x_shape):
    """Gradient w.r.t. the input of _raw_conv, through the forward kernels with re-packed weights:
    stride-1 conv -> conv with flipped taps and swapped channels; strided conv -> transposed conv;
    transposed conv -> strided conv (both with swapped channels, same tap order)."""
    swapped = wp.permute(0, 1, 2, 4, 3)
    if transposed:
        gx = _raw_conv(g, swapped.contiguous(), False, stride)
    elif tuple(stride) == (1, 1, 1):
        gx = _raw_conv(g, swapped.flip(0, 1, 2).contiguous(), False, stride)
    else:
        if stride[1] != 2 or stride[2] != 2:
            raise NotImplementedError("conv stride %s has no data-gradient kernel" % (tuple(stride),))
        gx = _raw_conv(g, swapped.contiguous(), True, stride)
    if tuple(gx.shape) != tuple(x_shape):
        raise RuntimeError("data gradient %s does not match the input %s (odd input size under a strided conv?)"
                           % (tuple(gx.shape), tuple(x_shape)))
    return gx


def _conv_wgrad(g, x, wp_shape, transposed, stride):
    kd, khw = wp_shape[0], wp_shape[1]
    if transposed:
        return conv_wgrad(x, g, kd, khw, stride[0], 2, small_is_cout=False)
    return conv_wgrad(g, x, kd, khw, stride[0], stride[1], small_is_cout=True)


def _group(bn):
    """Process group of a SyncBatchNorm layer (``convert_sync_batchnorm(module, process_group)``; None = default group)."""
    return getattr(bn, "process_group", None)


def _world(bn):
    """Ranks that share this layer's batch statistics.  Every rank must contribute the same number of elements
    (one reference view of the same crop size per rank, as the reference's DDP training does: train.py:46,138-139);
    the statistics are normalised by m * world."""
    if isinstance(bn, nn.SyncBatchNorm) and dist.is_available() and dist.is_initialized():
        return dist.get_world_size(_group(bn))
    return 1


class _ConvBnAct(torch.autograd.Function):
    """One Conv3d / Deconv3d / ConvBnReLU block in training (models/module.py:83-197)."""

    @staticmethod
    def forward(ctx, x, weight, gamma, beta, skip, bn, transposed, stride, relu):
        wp = _pack(weight, transposed)
        conv = _raw_conv(x, wp, transposed, stride)
        c = conv.shape[-1]
        m = conv.numel() // c
        world = 1
        if bn.training:
            if bn.momentum is None:
                raise NotImplementedError("BatchNorm with cumulative moving average (momentum=None) is not built")
            world = _world(bn)
            sums, replicas = _reduction_buffer(c, x.device), BN_REPLICAS
            _call("mvs_bn_stats", _p(conv), _p(sums), m, c)
            if world > 1:                      # SyncBatchNorm: same spatial size on every rank (DDP)
                _call("mvs_bn_collapse", _p(sums), c)
                sums, replicas = sums[:2 * c], 1
                dist.all_reduce(sums, group=_group(bn))
            mean_invstd = torch.empty(2 * c, device=x.device, dtype=torch.float32)
            track = bn.track_running_stats and bn.running_mean is not None
            _call("mvs_bn_finalize", _p(sums), replicas, float(m * world), float(bn.eps), float(bn.momentum), _p(mean_invstd),
                  _p(bn.running_mean if track else None), _p(bn.running_var if track else None), c)
            if track:
                # the kernel wrote through raw pointers: tell torch (the eval path's fold cache keys on versions)
                torch.autograd.graph.increment_version(bn.running_mean)
                torch.autograd.graph.increment_version(bn.running_var)
                if bn.num_batches_tracked is not None:
                    bn.num_batches_tracked += 1
        else:                                   # frozen BN inside a training graph: running statistics
            mean_invstd = torch.cat([bn.running_mean.float(), torch.rsqrt(bn.running_var.float() + bn.eps)]).contiguous()
        y = torch.empty_like(conv)
        _lib.require_cuda(conv, mean_invstd, gamma, beta, skip)
        _call("mvs_bn_act_fwd", _p(conv), _p(mean_invstd), _p(gamma), _p(beta), _p(skip), _p(y), m, c, 1 if relu else 0)
        ctx.save_for_backward(x, wp, conv, mean_invstd, gamma, beta)
        ctx.cfg = (transposed, tuple(stride), relu, bool(bn.training), world, skip is not None)
        ctx.group = _group(bn) if world > 1 else None
        return y

    @staticmethod
    def backward(ctx, gy):
        x, wp, conv, mean_invstd, gamma, beta = ctx.saved_tensors
        transposed, stride, relu, batch_stats, world, has_skip = ctx.cfg
        gy = gy.contiguous()
        _lib.require_cuda(gy)
        c = conv.shape[-1]
        m = conv.numel() // c
        sums = _reduction_buffer(c, gy.device)
        _call("mvs_bn_act_bwd_reduce", _p(gy), _p(conv), _p(mean_invstd), _p(gamma), _p(beta), _p(sums), m, c,
              1 if relu else 0)
        _call("mvs_bn_collapse", _p(sums), c)
        sums = sums[:2 * c]
        ggamma, gbeta = sums[:c].float(), sums[c:].float()          # this rank's share (DDP averages them)
        if batch_stats:
            if world > 1:
                sums = sums.clone()
                dist.all_reduce(sums, group=ctx.group)
            mean_terms = sums
        else:
            mean_terms = torch.zeros_like(sums)                      # statistics were constants
        g = torch.empty_like(conv)
        _call("mvs_bn_act_bwd_apply", _p(gy), _p(conv), _p(mean_invstd), _p(gamma), _p(beta), _p(mean_terms),
              float(m * world), _p(g), m, c, 1 if relu else 0)
        gweight = _unpack(_conv_wgrad(g, x, wp.shape, transposed, stride), transposed) if ctx.needs_input_grad[1] else None
        gx = _conv_dgrad(g, wp, transposed, stride, x.shape) if ctx.needs_input_grad[0] else None
        return (gx, gweight, ggamma if ctx.needs_input_grad[2] else None, gbeta if ctx.needs_input_grad[3] else None,
                gy if has_skip and ctx.needs_input_grad[4] else None, None, None, None, None)


def conv_bn_act(x, conv, bn, skip, transposed, stride, relu=True):
    """x channels-last [B,D,H,W,Cin]; ``conv`` an nn.Conv3d / nn.ConvTranspose3d / nn.Conv2d without
    bias, ``bn`` its BatchNorm module (statistics buffers are updated in place as nn.BatchNorm does)."""
    if conv.bias is not None:
        raise NotImplementedError("conv bias in front of a BatchNorm is not built for training")
    weight = conv.weight
    if weight.dim() == 4:                                   # Conv2d -> depth-1 3D kernel
        weight = weight.unsqueeze(2)
    gamma = bn.weight if bn.weight is not None else torch.ones(weight.shape[1 if transposed else 0], device=x.device)
    beta = bn.bias if bn.bias is not None else torch.zeros_like(gamma)
    return _ConvBnAct.apply(x.contiguous(), weight.float(), gamma.float(), beta.float(),
                            None if skip is None else skip.contiguous(), bn, transposed, tuple(stride), relu)


# ------------------------------------------------------------------------------------------------
# thin convolution with bias (vis head 8->1 + sigmoid, `prob` 8->1)
# ------------------------------------------------------------------------------------------------


class _ThinConv(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, act):
        cout, cin, kd, khw, _ = weight.shape
        w_taps = weight.permute(2, 3, 4, 1, 0).contiguous().view(kd * khw * khw, cin, cout)
        y = thin_conv(x, w_taps, bias, kd, khw, act)
        ctx.save_for_backward(x, w_taps, y if act == 2 else None)
        ctx.cfg = (kd, khw, act, bias is not None)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, w_taps, y = ctx.saved_tensors
        kd, khw, act, has_bias = ctx.cfg
        g = gy.contiguous()
        _lib.require_cuda(g)
        if act == 2:
            gs = torch.empty_like(g)
            _call("mvs_sigmoid_bwd", _p(g), _p(y), _p(gs), g.numel())
            g = gs
        elif act != 0:
            raise NotImplementedError("thin conv backward through ReLU is not built")
        _, cin, cout = w_taps.shape
        gbias = channel_sums(g, cout)[:cout].float() if has_bias and ctx.needs_input_grad[2] else None
        gweight = None
        if ctx.needs_input_grad[1]:
            dwp = conv_wgrad(g, x, kd, khw, 1, 1, small_is_cout=True)            # [kd,k,k,Cin,Cout]
            gweight = _unpack(dwp, False)
        gx = None
        if ctx.needs_input_grad[0]:
            wd = w_taps.view(kd, khw, khw, cin, cout).flip(0, 1, 2).permute(0, 1, 2, 4, 3).contiguous()
            gx = thin_conv(g, wd.view(-1, cout, cin), None, kd, khw, 0)
        return gx, gweight, gbias, None


def thin_conv_module(x, conv, act=0):
    """x channels-last [B,D,H,W,Cin] through an nn.Conv3d / nn.Conv2d (stride 1, 'same' padding) ->
    [B,D,H,W,Cout]; act 0 none, 2 sigmoid."""
    weight = conv.weight
    if weight.dim() == 4:
        weight = weight.unsqueeze(2)
    k = tuple(weight.shape[2:])
    if k[1] != k[2] or k[0] not in (1, 3) or k[1] not in (1, 3) or any(s != 1 for s in conv.stride):
        raise NotImplementedError("thin conv geometry %s / stride %s is not built" % (k, tuple(conv.stride)))
    bias = conv.bias.float() if conv.bias is not None else None
    return _ThinConv.apply(x.contiguous(), weight.float(), bias, act)


# ------------------------------------------------------------------------------------------------
# head: softmax (differentiable), argmax depth and confidence (not)      mvsformer_model.py:110-125
# ------------------------------------------------------------------------------------------------


class _Head(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pre, depth_values, tmp):
        prob, depth, conf = engine.regression_head(pre, depth_values, tmp, True)
        ctx.save_for_backward(prob)
        ctx.mark_non_differentiable(depth, conf)
        return prob, depth, conf

    @staticmethod
    def backward(ctx, gprob, _gdepth, _gconf):
        (prob,) = ctx.saved_tensors
        gprob = gprob.contiguous()
        _lib.require_cuda(gprob)
        b, d, h, w = prob.shape
        gpre = torch.empty_like(prob)
        _call("mvs_softmax_bwd", _p(gprob), _p(prob), _p(gpre), b, d, h, w)
        return gpre, None, None


def train_head(pre, depth_values, tmp):
    return _Head.apply(pre, depth_values, float(tmp))


class _DepthRegression(torch.autograd.Function):
    """depth_regression(p, depth_values) (models/module.py:597-603), differentiable w.r.t. p."""

    @staticmethod
    def forward(ctx, p, depth_values):
        ctx.save_for_backward(depth_values)
        ctx.p_shape = tuple(p.shape)
        return engine.depth_regression(p, depth_values)

    @staticmethod
    def backward(ctx, gdepth):
        (depth_values,) = ctx.saved_tensors
        gdepth = gdepth.contiguous()
        dv = depth_values.float().contiguous()
        _lib.require_cuda(gdepth, dv)
        b, d, h, w = ctx.p_shape
        gp = torch.empty(b, d, h, w, device=gdepth.device, dtype=torch.float32)
        _call("mvs_depth_regression_bwd", _p(gdepth), _p(dv), 1 if dv.dim() == 4 else 0, _p(gp), b, d, h, w)
        return gp, None


def depth_regression(p, depth_values):
    return _DepthRegression.apply(engine._f32(p).contiguous(), depth_values)


def mixup_head(prob, depth_values):
    """depth_type 'mixup_ce' (models/mvsformer_model.py:126-136) -> (depth, confidence), not differentiable (the
    reference's losses for this head use prob_volume_pre only, models/losses.py:360-398)."""
    prob, dv = prob.detach().float().contiguous(), depth_values.float().contiguous()
    _lib.require_cuda(prob, dv)
    b, d, h, w = prob.shape
    depth = torch.empty(b, h, w, device=prob.device, dtype=torch.float32)
    conf = torch.empty(b, h, w, device=prob.device, dtype=torch.float32)
    _call("mvs_mixup_head", _p(prob), _p(dv), _p(depth), _p(conf), b, d, h, w)
    return depth, conf


# ------------------------------------------------------------------------------------------------
# fusion_type 'epipole' / 'epipoleV2'                                  models/mvsformer_model.py:92-104
# ------------------------------------------------------------------------------------------------


def proj_mask(relproj, depth_values):
    """[B,N,12], [B,D,H,W] -> float 1/0 [B,N,D,H,W]: sample outside the source image or behind it (warping.py:99-103)."""
    dv = depth_values.float().contiguous()
    _lib.require_cuda(relproj, dv)
    b, n = relproj.shape[:2]
    d, h, w = dv.shape[1:]
    mask = torch.empty(b, n, d, h, w, device=dv.device, dtype=torch.float32)
    _call("mvs_proj_mask", _p(relproj), _p(dv), _p(mask), b, n, d, h, w)
    return mask


class _EpipoleAggregate(torch.autograd.Function):
    """volume = sum_v corr_v w_v / (sum_v w_v + 1e-6) with per-hypothesis softmax weights; differentiable w.r.t. corr
    (both through the product and through the weights) and the temperature."""

    @staticmethod
    def forward(ctx, corr, temperature, mask, norm, clamp):
        t_eff = float(temperature)
        if clamp is not None:
            t_eff = min(max(t_eff, clamp[0]), clamp[1])
        _lib.require_cuda(corr, mask)
        b, n, d, h, w, g = corr.shape
        stats = torch.empty(b, n, h, w, 2, device=corr.device, dtype=torch.float32)
        volume = torch.empty(b, d, h, w, g, device=corr.device, dtype=torch.float32)
        wsum = torch.empty(b, d, h, w, device=corr.device, dtype=torch.float32)
        _call("mvs_epipole_aggregate_fwd", _p(corr), _p(mask), t_eff, float(norm), _p(stats), _p(volume), _p(wsum), b, n, g, d, h, w)
        ctx.save_for_backward(corr, mask, stats, volume, wsum)
        inside = clamp is None or (clamp[0] <= float(temperature) <= clamp[1])
        ctx.cfg = (t_eff, float(norm), inside)
        return volume

    @staticmethod
    def backward(ctx, gvol):
        corr, mask, stats, volume, wsum = ctx.saved_tensors
        t_eff, norm, inside = ctx.cfg
        gvol = gvol.contiguous()
        _lib.require_cuda(gvol)
        b, n, d, h, w, g = corr.shape
        gcorr = torch.empty_like(corr)
        want_t = ctx.needs_input_grad[1]
        gtemp = torch.zeros(32, device=corr.device, dtype=torch.float32) if want_t else None
        _call("mvs_epipole_aggregate_bwd", _p(gvol), _p(corr), _p(mask), _p(stats), _p(volume), _p(wsum), t_eff, norm, _p(gcorr),
              _p(gtemp), b, n, g, d, h, w)
        gt = None
        if want_t:
            gt = gtemp.sum().reshape(()) if inside else torch.zeros((), device=corr.device)      # clamp passes gradient inside its range
        return gcorr, gt, None, None, None


def epipole_aggregate(corr, temperature, mask, norm, clamp=None):
    """corr [B,N,D,H,W,G]; ``temperature`` a python float ('epipole') or a 0-d tensor / nn.Parameter ('epipoleV2');
    ``mask`` [B,N,D,H,W] or None; ``clamp`` = (lo, hi) applied to the temperature (V2: 0.1, 10)."""
    if not torch.is_tensor(temperature):
        temperature = torch.tensor(float(temperature), device=corr.device)
    return _EpipoleAggregate.apply(corr.contiguous(), temperature, mask, float(norm), clamp)
