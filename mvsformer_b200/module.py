"""Drop-in mirror of the hot-path half of the reference's ``models/module.py``.

Same class / function names, constructor arguments, ``state_dict`` keys and shapes
(``load_state_dict(strict=True)`` of a reference checkpoint works), same call signatures — the
arithmetic runs in libmvs_b200.so on channels-last activations.  Reference lines:

* ``Conv3d`` / ``Deconv3d`` blocks      models/module.py:83-159
* ``ConvBnReLU``                         models/module.py:168-197
* ``CostRegNet``                         models/module.py:469-505
* ``CostRegNet2D``                       models/module.py:508-547
* ``CostRegNet3D``                       models/module.py:550-594
* ``depth_regression`` ... ``schedule_range``   models/module.py:597-699

Eval-mode BatchNorm is folded into the convolution (scale into the packed weights, shift as a
per-channel bias) once per parameter version.  In training (``module.training``) every block
runs through ``mvsformer_b200.autograd``: raw FP32 convolution -> batch-statistics BatchNorm (running
statistics updated in place) -> ReLU (+ skip), with a hand-written backward.
"""
import numpy as np
import torch
import torch.nn as nn

from . import autograd, config, engine


def _versions(*tensors):
    return tuple((t.data_ptr(), t._version) for t in tensors if t is not None)


def _bn_scale_shift(bn):
    inv = torch.rsqrt(bn.running_var.float() + bn.eps)
    gamma = bn.weight.float() if bn.weight is not None else torch.ones_like(inv)
    beta = bn.bias.float() if bn.bias is not None else torch.zeros_like(inv)
    scale = gamma * inv
    return scale, beta - bn.running_mean.float() * scale


class _FoldCache:
    """Folded, packed parameters of one conv(+BN) block, rebuilt when a tensor changes."""

    def __init__(self):
        self.key = None
        self.value = None
        self.derived = {}

    def get(self, tensors, build):
        key = _versions(*tensors)
        if key != self.key:
            with torch.no_grad():
                self.value = build()
            self.key = key
            self.derived = {}
        return self.value

    def get_derived(self, name, build):
        """Secondary packing (e.g. tensor-core operand order) of the current folded value."""
        if name not in self.derived:
            with torch.no_grad():
                self.derived[name] = build(self.value)
        return self.derived[name]


_TC_CONV_TILES = {(8, 16), (16, 16), (16, 32), (32, 16), (32, 32), (32, 64)}      # (channel slice, N tile) built in conv3d_tc.cu
_TC_DECONV_TILES = {(16, 16), (32, 16), (32, 32)}


def _tc_eligible(cin, cout, transposed):
    """True when the tcgen05 kernels have an instantiation for this layer shape."""
    if config.conv_precision() == "fp32":
        return False
    if not (cin in (8, 16) or cin % 32 == 0) or cout % 8 or cout > 64:
        return False
    x3 = config.conv_precision() == "tf32x3"
    key = (engine.tc_channel_slice(cin), engine.tc_n_tile(cout, x3, transposed))
    return key in (_TC_DECONV_TILES if transposed else _TC_CONV_TILES)


def _run_conv(cache, x, w, shift, skip, stride, relu):
    kd, cin, cout = w.shape[0], w.shape[3], w.shape[4]
    if (config.conv_precision() == "tf32" and config.conv_tma() and kd == 3 and stride[0] == 1 and stride[1] == stride[2]
            and engine.tma_supported(cin, cout, x.shape[1], kd, stride[1] == 2)):
        # round-2 persistent TMA-fed tcgen05 kernels (csrc/conv3d_tma.cu)
        mode = engine.TMA_S2 if stride[1] == 2 else engine.TMA_S1
        wt, nt = cache.get_derived("tma_m%d" % mode, lambda v: engine.pack_tma_weights(v[0], mode))
        return engine.conv3d_tma(x, wt, nt, cout, kd, shift, skip, relu, mode)
    if stride[1] == stride[2] and _tc_eligible(cin, cout, False):
        x3 = config.conv_precision() == "tf32x3"
        hi, lo, nt = cache.get_derived("tc_x3" if x3 else "tc", lambda v: engine.pack_tc_weights(v[0], x3))
        return engine.conv3d_tc(x, hi, lo, nt, cout, kd, shift, skip, stride, relu)
    return engine.conv3d_cl(x, w, shift, skip, stride, relu)


def _run_deconv(cache, x, w, shift, skip, sd, relu):
    kd, cin, cout = w.shape[0], w.shape[3], w.shape[4]
    if (config.conv_precision() == "tf32" and config.conv_tma() and kd == 3 and sd == 1
            and engine.tma_supported(cin, cout, x.shape[1], kd, transposed=True)):
        wt, nt = cache.get_derived("tma_m2", lambda v: engine.pack_tma_weights(v[0], engine.TMA_DECONV))
        return engine.conv3d_tma(x, wt, nt, cout, kd, shift, skip, relu, engine.TMA_DECONV)
    if _tc_eligible(cin, cout, True):
        x3 = config.conv_precision() == "tf32x3"
        hi, lo, nt = cache.get_derived("tcd_x3" if x3 else "tcd", lambda v: engine.pack_tc_deconv_weights(v[0], x3))
        return engine.deconv3d_tc(x, hi, lo, nt, cout, kd, shift, skip, sd, relu)
    return engine.deconv3d_cl(x, w, shift, skip, sd, relu)


def _pack_conv(conv, bn, transposed):
    """-> (w_packed [kd,kh,kw,Cin,Cout] fp32 contiguous, shift [Cout] or None)."""
    w = conv.weight.detach().float()
    if transposed:                       # [Cin,Cout,kd,kh,kw]
        w = w.permute(2, 3, 4, 0, 1)
    else:                                # [Cout,Cin,kd,kh,kw]
        w = w.permute(2, 3, 4, 1, 0)
    shift = None
    if bn is not None:
        scale, shift = _bn_scale_shift(bn)
        w = w * scale.view(1, 1, 1, 1, -1)
        shift = shift.contiguous()
    if conv.bias is not None:
        b = conv.bias.detach().float()
        shift = b if shift is None else shift + b * (scale if bn is not None else 1.0)
        shift = shift.contiguous()
    return w.contiguous(), shift


def _triple(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (v, v, v)


def _train_block(mod, conv, bn, x, skip, transposed, stride, relu):
    """Training forward of one conv(+BN+ReLU) block (autograd.conv_bn_act)."""
    if bn is None:
        raise NotImplementedError("%s without BatchNorm is not built for training" % type(mod).__name__)
    return autograd.conv_bn_act(x, conv, bn, skip, transposed, stride, relu)


class Conv3d(nn.Module):
    """conv3d(bias = not bn) -> BatchNorm3d -> ReLU block (models/module.py:83-117)."""

    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, relu=True, bn=True, bn_momentum=0.1,
                 init_method="xavier", **kwargs):
        super().__init__()
        self.out_channels = out_channels
        self.kernel_size = kernel_size
        self.conv = nn.Conv3d(in_channels, out_channels, kernel_size, stride=stride, bias=(not bn), **kwargs)
        self.bn = nn.BatchNorm3d(out_channels, momentum=bn_momentum) if bn else None
        self.relu = relu
        self._fold = _FoldCache()

    def _check_geometry(self):
        k, p = _triple(self.conv.kernel_size), _triple(self.conv.padding)
        if k[1:] != (3, 3) or k[0] not in (1, 3) or p != (k[0] // 2, 1, 1) or _triple(self.conv.dilation) != (1, 1, 1):
            raise NotImplementedError("Conv3d: only (1|3,3,3) kernels with 'same' padding are built, got k=%s p=%s" % (k, p))

    def packed(self):
        tensors = [self.conv.weight, self.conv.bias]
        if self.bn is not None:
            tensors += [self.bn.weight, self.bn.bias, self.bn.running_mean, self.bn.running_var]
        return self._fold.get(tensors, lambda: _pack_conv(self.conv, self.bn, False))

    def forward_cl(self, x, skip=None):
        """Channels-last forward: x [B,D,H,W,Cin]; ``skip`` is added after the activation."""
        self._check_geometry()
        if self.training:
            return _train_block(self, self.conv, self.bn, x, skip, False, _triple(self.conv.stride), self.relu)
        w, shift = self.packed()
        return _run_conv(self._fold, x, w, shift, skip, _triple(self.conv.stride), self.relu)

    def forward(self, x):
        if self.training:                  # differentiable layout change (torch owns the permutes)
            return self.forward_cl(x.float().permute(0, 2, 3, 4, 1).contiguous()).permute(0, 4, 1, 2, 3)
        return engine.cl_to_ncdhw(self.forward_cl(engine.ncdhw_to_cl(x, config.conv_precision() == "tf32")))

    def init_weights(self, init_method):
        _init_uniform(self.conv, init_method)
        if self.bn is not None:
            nn.init.ones_(self.bn.weight)
            nn.init.zeros_(self.bn.bias)


class Deconv3d(nn.Module):
    """ConvTranspose3d(bias = not bn) -> BatchNorm3d -> ReLU block (models/module.py:126-159)."""

    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, relu=True, bn=True, bn_momentum=0.1,
                 init_method="xavier", **kwargs):
        super().__init__()
        self.out_channels = out_channels
        self.conv = nn.ConvTranspose3d(in_channels, out_channels, kernel_size, stride=stride, bias=(not bn), **kwargs)
        self.bn = nn.BatchNorm3d(out_channels, momentum=bn_momentum) if bn else None
        self.relu = relu
        self._fold = _FoldCache()

    def packed(self):
        tensors = [self.conv.weight, self.conv.bias]
        if self.bn is not None:
            tensors += [self.bn.weight, self.bn.bias, self.bn.running_mean, self.bn.running_var]
        return self._fold.get(tensors, lambda: _pack_conv(self.conv, self.bn, True))

    def forward_cl(self, x, skip=None):
        sd = _deconv_depth_stride(self.conv)
        if self.training:
            return _train_block(self, self.conv, self.bn, x, skip, True, (sd, 2, 2), self.relu)
        w, shift = self.packed()
        return _run_deconv(self._fold, x, w, shift, skip, sd, self.relu)

    def forward(self, x):
        if self.training:                  # differentiable layout change (torch owns the permutes)
            return self.forward_cl(x.float().permute(0, 2, 3, 4, 1).contiguous()).permute(0, 4, 1, 2, 3)
        return engine.cl_to_ncdhw(self.forward_cl(engine.ncdhw_to_cl(x, config.conv_precision() == "tf32")))

    def init_weights(self, init_method):
        _init_uniform(self.conv, init_method)
        if self.bn is not None:
            nn.init.ones_(self.bn.weight)
            nn.init.zeros_(self.bn.bias)


def _deconv_depth_stride(conv):
    """Validates the transposed-conv geometry the kernels implement and returns the depth stride."""
    k, s, p, op = (_triple(conv.kernel_size), _triple(conv.stride), _triple(conv.padding), _triple(conv.output_padding))
    ok = (k[1:] == (3, 3) and k[0] in (1, 3) and s[1:] == (2, 2) and s[0] in (1, 2)
          and p == (k[0] // 2, 1, 1) and op == (s[0] - 1, 1, 1) and not (k[0] == 1 and s[0] != 1))
    if not ok:
        raise NotImplementedError("ConvTranspose3d geometry k=%s s=%s p=%s op=%s is not built" % (k, s, p, op))
    return s[0]


def _init_uniform(module, init_method):
    if module.weight is not None:
        if init_method == "kaiming":
            nn.init.kaiming_uniform_(module.weight)
        elif init_method == "xavier":
            nn.init.xavier_uniform_(module.weight)


class ConvBnReLU(nn.Module):
    """2D conv + BN + ReLU parameter block (models/module.py:168-197).  On the hot path it only
    occurs inside ``StageNet.vis``, which executes as ONE fused kernel (vis_net.cu); the block
    itself holds the parameters under the reference's key names."""

    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, pad=1, dilation=1):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride=stride, padding=pad, dilation=dilation,
                              bias=False)
        self.bn = nn.BatchNorm2d(out_channels)

    def forward_cl(self, x):
        """x channels-last [B,1,H,W,Cin] (a depth-1 volume) -> [B,1,H,W,Cout]; FP32 conv, BatchNorm with batch
        statistics in training / running statistics in eval, ReLU (autograd.conv_bn_act)."""
        k, st, pd, dl = (tuple(self.conv.kernel_size), tuple(self.conv.stride), tuple(self.conv.padding),
                         tuple(self.conv.dilation))
        if k != (3, 3) or st != (1, 1) or pd != (1, 1) or dl != (1, 1):
            raise NotImplementedError("ConvBnReLU: only 3x3 / stride 1 / pad 1 is built, got k=%s s=%s p=%s d=%s" % (k, st, pd, dl))
        return autograd.conv_bn_act(x, self.conv, self.bn, None, False, (1, 1, 1), True)

    def forward(self, x):
        """Standalone use (NCHW in / out).  Inside ``StageNet.vis`` the eval path is one fused kernel instead."""
        y = self.forward_cl(x.float().permute(0, 2, 3, 1).unsqueeze(1).contiguous())
        return y.squeeze(1).permute(0, 3, 1, 2)


class _SeqDeconv(nn.Sequential):
    """nn.Sequential(ConvTranspose3d(bias=False), BatchNorm3d, ReLU) with the reference's key
    layout (``conv7.0.weight``, ``conv7.1.*``; models/module.py:562-575) and a fused forward."""

    def __init__(self, in_channels, out_channels, kernel_size, stride, padding, output_padding):
        super().__init__(
            nn.ConvTranspose3d(in_channels, out_channels, kernel_size=kernel_size, padding=padding,
                               output_padding=output_padding, stride=stride, bias=False),
            nn.BatchNorm3d(out_channels),
            nn.ReLU(inplace=True))
        self._fold = _FoldCache()

    def packed(self):
        conv, bn = self[0], self[1]
        return self._fold.get([conv.weight, bn.weight, bn.bias, bn.running_mean, bn.running_var],
                              lambda: _pack_conv(conv, bn, True))

    def forward_cl(self, x, skip=None):
        sd = _deconv_depth_stride(self[0])
        if self.training:
            return _train_block(self, self[0], self[1], x, skip, True, (sd, 2, 2), True)
        w, shift = self.packed()
        return _run_deconv(self._fold, x, w, shift, skip, sd, True)

    def forward(self, x):
        if self.training:                  # differentiable layout change (torch owns the permutes)
            return self.forward_cl(x.float().permute(0, 2, 3, 4, 1).contiguous()).permute(0, 4, 1, 2, 3)
        return engine.cl_to_ncdhw(self.forward_cl(engine.ncdhw_to_cl(x, config.conv_precision() == "tf32")))


class _ProbCache(_FoldCache):
    pass


def _prob_host(conv):
    """8->1 `prob` conv as host arrays for mvs_prob_conv_cl: w [k^3][Cin], bias [1] or None."""
    w = conv.weight.detach().float()                       # [1, Cin, k, k, k]
    w_host = w[0].permute(1, 2, 3, 0).contiguous().cpu().numpy().astype(np.float32)
    b_host = conv.bias.detach().float().cpu().numpy().astype(np.float32) if conv.bias is not None else None
    return np.ascontiguousarray(w_host.reshape(-1, w.shape[1])), b_host


class _RegNetBase(nn.Module):
    """Shared U-Net driver: 3 strided levels down, 3 transposed levels up with skip adds after
    the ReLU (models/module.py:495-505 / :584-594)."""

    def _unet_cl(self, x, with_prob=False):
        """-> the U-Net's output [B,D,H,W,base]; ``with_prob``: -> prob_volume_pre [B,D,H,W] when the last transposed layer can
        take the ``prob`` conv into its epilogue (``_prob_fusable``), else the output as before."""
        c0 = x
        c2 = self.conv2.forward_cl(self.conv1.forward_cl(c0))
        c4 = self.conv4.forward_cl(self.conv3.forward_cl(c2))
        y = self.conv6.forward_cl(self.conv5.forward_cl(c4))
        y = self.conv7.forward_cl(y, skip=c4)
        y = self.conv9.forward_cl(y, skip=c2)
        if not isinstance(self.inner, nn.Identity):
            # 1x1x1 projection of the input skip when in_channels != base_channels (module.py:486-489, :502)
            c0 = autograd.thin_conv_module(c0, self.inner)
        if with_prob and self._prob_fusable(y):
            w_host, b_host = self._prob_weights()
            wd, shift = self.conv11.packed()
            wt, _ = self.conv11._fold.get_derived("tma_m2", lambda v: engine.pack_tma_weights(v[0], engine.TMA_DECONV))
            return engine.conv3d_tma_prob(y, wt, wd.shape[0], shift, c0, w_host, float(b_host[0]) if b_host is not None else 0.0), True
        y = self.conv11.forward_cl(y, skip=c0)
        return (y, False) if with_prob else y

    def _prob_fusable(self, y):
        """Eval, TF32 mode, the persistent TMA kernels on, a 3x3x3 16 -> 8 transposed last layer (depth stride 1: the layer
        _run_deconv sends to mvs_conv3d_tma) followed by a 1x1x1 ``prob`` conv 8 -> 1: CostRegNet3D as the shipped configs
        build it (stages 3-4)."""
        conv11, prob = self.conv11, self.prob
        if self.training or not config.prob_fused() or config.conv_precision() != "tf32" or not config.conv_tma():
            return False
        if (not isinstance(conv11, (_SeqDeconv, Deconv3d)) or tuple(prob.kernel_size) != (1, 1, 1) or prob.in_channels != 8
                or prob.out_channels != 1):
            return False
        conv = conv11[0] if isinstance(conv11, _SeqDeconv) else conv11.conv
        if isinstance(conv11, Deconv3d) and (conv11.bn is None or not conv11.relu):
            return False
        kd = _triple(conv.kernel_size)[0]
        return (conv.in_channels == 16 and conv.out_channels == 8 and _deconv_depth_stride(conv) == 1 and kd == 3
                and engine.tma_supported(16, 8, y.shape[1], kd, transposed=True))

    def _prob_weights(self):
        conv = self.prob
        if not hasattr(self, "_prob_cache"):
            object.__setattr__(self, "_prob_cache", _ProbCache())
        return self._prob_cache.get([conv.weight, conv.bias], lambda: _prob_host(conv))

    def _prob_cl(self, y):
        conv = self.prob
        if self.training:
            return autograd.thin_conv_module(y, conv).squeeze(-1)
        w_host, b_host = self._prob_weights()
        return engine.prob_conv_cl(y, w_host, b_host, conv.kernel_size[0])

    def forward_cl(self, x):
        """x: channels-last cost volume [B,D,H,W,Cin] -> prob_volume_pre [B,D,H,W]."""
        b, d, h, w, _ = x.shape
        sd = self._depth_stride
        if h % 8 or w % 8 or (sd == 2 and d % 8):
            # the reference fails at the first skip add with torch's size-mismatch RuntimeError
            raise RuntimeError("The size of tensor a must match the size of tensor b: cost volume [D=%d,H=%d,W=%d] "
                               "is not divisible by the U-Net's total stride 8" % (d, h, w))
        if getattr(self, "last_layer", True):
            y, done = self._unet_cl(x, with_prob=True)
            return y if done else self._prob_cl(y)
        return self._unet_cl(x)

    def forward(self, x):
        if self.training:
            out = self.forward_cl(x.float().permute(0, 2, 3, 4, 1).contiguous())
            return out.unsqueeze(1) if out.dim() == 4 else out.permute(0, 4, 1, 2, 3)
        out = self.forward_cl(engine.ncdhw_to_cl(x, config.conv_precision() == "tf32"))
        if out.dim() == 4:
            return out.unsqueeze(1)
        return engine.cl_to_ncdhw(out)


class CostRegNet(_RegNetBase):
    """models/module.py:469-505: stride 2 in D,H,W; ``prob`` = 3x3x3 conv 8->1 without bias."""
    _depth_stride = 2

    def __init__(self, in_channels, base_channels, last_layer=True):
        super().__init__()
        self.last_layer = last_layer
        self.conv1 = Conv3d(in_channels, base_channels * 2, stride=2, padding=1)
        self.conv2 = Conv3d(base_channels * 2, base_channels * 2, padding=1)
        self.conv3 = Conv3d(base_channels * 2, base_channels * 4, stride=2, padding=1)
        self.conv4 = Conv3d(base_channels * 4, base_channels * 4, padding=1)
        self.conv5 = Conv3d(base_channels * 4, base_channels * 8, stride=2, padding=1)
        self.conv6 = Conv3d(base_channels * 8, base_channels * 8, padding=1)
        self.conv7 = Deconv3d(base_channels * 8, base_channels * 4, stride=2, padding=1, output_padding=1)
        self.conv9 = Deconv3d(base_channels * 4, base_channels * 2, stride=2, padding=1, output_padding=1)
        self.conv11 = Deconv3d(base_channels * 2, base_channels * 1, stride=2, padding=1, output_padding=1)
        self.inner = nn.Conv3d(in_channels, base_channels, 1, 1) if in_channels != base_channels else nn.Identity()
        if self.last_layer:
            self.prob = nn.Conv3d(base_channels, 1, 3, stride=1, padding=1, bias=False)


class CostRegNet3D(_RegNetBase):
    """models/module.py:550-594: stride (1,2,2); ``prob`` = 1x1x1 conv 8->1 with bias."""
    _depth_stride = 1

    def __init__(self, in_channels, base_channel=8):
        super().__init__()
        bc = base_channel
        self.conv1 = Conv3d(in_channels, bc * 2, kernel_size=3, stride=(1, 2, 2), padding=1)
        self.conv2 = Conv3d(bc * 2, bc * 2, padding=1)
        self.conv3 = Conv3d(bc * 2, bc * 4, kernel_size=3, stride=(1, 2, 2), padding=1)
        self.conv4 = Conv3d(bc * 4, bc * 4, padding=1)
        self.conv5 = Conv3d(bc * 4, bc * 8, kernel_size=3, stride=(1, 2, 2), padding=1)
        self.conv6 = Conv3d(bc * 8, bc * 8, padding=1)
        self.conv7 = _SeqDeconv(bc * 8, bc * 4, 3, (1, 2, 2), 1, (0, 1, 1))
        self.conv9 = _SeqDeconv(bc * 4, bc * 2, 3, (1, 2, 2), 1, (0, 1, 1))
        self.conv11 = _SeqDeconv(bc * 2, bc, 3, (1, 2, 2), 1, (0, 1, 1))
        self.inner = nn.Conv3d(in_channels, bc, 1, 1) if in_channels != bc else nn.Identity()
        self.prob = nn.Conv3d(bc, 1, 1, stride=1, padding=0)


class CostRegNet2D(_RegNetBase):
    """models/module.py:508-547 (not used by the shipped configs): (1,3,3) kernels on the strided
    and transposed layers, 3x3x3 on the stride-1 layers; skip of the input without ``inner``."""
    _depth_stride = 1

    def __init__(self, in_channels, base_channel=8):
        super().__init__()
        bc = base_channel
        k2, s2, p2 = (1, 3, 3), (1, 2, 2), (0, 1, 1)
        self.conv1 = Conv3d(in_channels, bc * 2, kernel_size=k2, stride=s2, padding=p2)
        self.conv2 = Conv3d(bc * 2, bc * 2, padding=1)
        self.conv3 = Conv3d(bc * 2, bc * 4, kernel_size=k2, stride=s2, padding=p2)
        self.conv4 = Conv3d(bc * 4, bc * 4, padding=1)
        self.conv5 = Conv3d(bc * 4, bc * 8, kernel_size=k2, stride=s2, padding=p2)
        self.conv6 = Conv3d(bc * 8, bc * 8, padding=1)
        self.conv7 = _SeqDeconv(bc * 8, bc * 4, k2, s2, p2, (0, 1, 1))
        self.conv9 = _SeqDeconv(bc * 4, bc * 2, k2, s2, p2, (0, 1, 1))
        self.conv11 = _SeqDeconv(bc * 2, bc, k2, s2, p2, (0, 1, 1))
        self.inner = nn.Identity()
        self.prob = nn.Conv3d(bc, 1, 1, stride=1, padding=0)


# ------------------------------------------------------------------------------------------------
# regression / scheduling functions (models/module.py:597-699), same signatures
# ------------------------------------------------------------------------------------------------


def depth_regression(p, depth_values):
    """models/module.py:597-603."""
    return engine.depth_regression(p, depth_values)


def conf_regression(p, n=4):
    """models/module.py:606-619."""
    return engine.conf_regression(p, n)


def init_range(cur_depth, ndepths, device, dtype, H, W):
    """models/module.py:622-630 (``device`` / ``dtype`` kept for signature compatibility)."""
    return engine.init_range(cur_depth, ndepths, H, W, inverse=False)


def init_inverse_range(cur_depth, ndepths, device, dtype, H, W):
    """models/module.py:633-639."""
    return engine.init_range(cur_depth, ndepths, H, W, inverse=True)


def schedule_inverse_range(depth, depth_hypo, ndepths, split_itv, H, W):
    """models/module.py:642-653."""
    return engine.schedule_inverse_range(depth, depth_hypo, ndepths, split_itv, H, W)


def schedule_range(cur_depth, ndepth, depth_inteval_pixel, H, W):
    """models/module.py:687-699."""
    return engine.schedule_range(cur_depth, ndepth, depth_inteval_pixel, H, W)
