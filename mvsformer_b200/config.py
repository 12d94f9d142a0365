"""Run-time knobs of the engine.

``conv_precision`` selects the arithmetic of the 3D-CNN convolutions (mvs_conv3d_* / mvs_deconv3d_*):

* ``"tf32x3"`` (default) — tcgen05 tensor cores, each operand split into two TF32 terms and three
  MMAs per product (error ~2^-21: fp32-grade; keeps the fp32 parity gates of the tests).
* ``"tf32"``   — tcgen05 tensor cores, operands rounded to TF32 (what cuDNN does by default for
  the reference on a GPU, ``torch.backends.cudnn.allow_tf32 = True``).
* ``"fp32"``   — FP32 CUDA-core kernels (exact fp32 FMA chains).

Set with ``set_conv_precision()`` or the environment variable ``MVS_CONV_PRECISION``.

``cv_store`` (``MVS_CV_STORE``, default ON since round 2; ``0`` restores two sampling passes) — cost-volume build with
ONE sampling pass at the stages where the per-view group correlation is smaller than the warped tensor (C/G >= 2,
stages 1-3): pass A stores it, the aggregation streams over it (bit-identical volume; +2 x N x G x D x h x w x 4 B of
HBM traffic instead of a second warp).  Measured on B200 (profiles/r02_ab_variants.json): 7.87 -> 7.30 ms per
reference view at cfg 2, refined depth bit-identical.

``cv_layout`` (``MVS_CV_LAYOUT`` = ``cl`` (default) | ``nchw``) — which cost-volume kernels sample the features: the
round-2 channels-last kernels (csrc/cost_volume_cl.cu: one LDS.128 per 4-channel tap, conflict-free lane order, next
item's TMA tile in flight; features are re-laid out once per call by mvs_features_to_cl) or the round-1 channel-planar
TMA kernels (csrc/cost_volume_tma.cu), kept for A/B runs.

``vis_fused`` (``MVS_VIS_FUSED``, default 1) — the visibility net as one persistent kernel (csrc/vis_fused.cu, TF32 mode):
only the entropy map and the weight map touch HBM; ``0`` = the four round-1 kernels.

``conv_tma`` (``MVS_CONV_TMA``, default 1; ``0`` = round-1 kernels only) — depth-unstrided 3x3x3 layers (CostRegNet3D) through the
persistent, warp-specialised, TMA-fed tcgen05 kernels of csrc/conv3d_tma.cu (TF32 mode only).

``tcz_kzf`` (``MVS_TCZ_KZF``: ``0`` off, ``1`` default, ``2`` = also prefer the kz-fused kernel over the row-tiled one) — "fused N"
variants of the depth-unstrided tensor-core convolutions (mvs_conv3d_tcz_kzf, mvs_deconv3d_tcz_kzf, mvs_conv3d_tcr_khf):
one MMA of N = 3 x Cout-tile per slab / input row instead of three, i.e. about a third of the shared-memory
A-operand reads.  TF32 mode only.  Default 1 since round 2 (measured 7.87 -> 7.43 ms, refined depth rel-L1 2e-6 vs the unfused kernels;
level 2 was slower: 8.93 ms).

``train_conv`` (``MVS_TRAIN_CONV`` = ``fp32`` (default) | ``tf32x3`` | ``tf32``) — arithmetic of the training path's
forward and data-gradient convolutions: the FP32 CUDA-core kernels, or the tcgen05 kernels of the inference path
(``tf32x3`` keeps fp32-grade accuracy; the reference itself trains the regulariser under fp16 autocast).  Weight
gradients stay FP32.  Opt-in until timed on the GPU.
"""
import os

_VALID = ("tf32x3", "tf32", "fp32")
_state = {"conv_precision": os.environ.get("MVS_CONV_PRECISION", "tf32x3"),
          "cv_store": os.environ.get("MVS_CV_STORE", "1") not in ("", "0"),
          "cv_layout": os.environ.get("MVS_CV_LAYOUT", "cl"),
          "vis_fused": os.environ.get("MVS_VIS_FUSED", "1") not in ("", "0"),
          "conv_tma": os.environ.get("MVS_CONV_TMA", "1") not in ("", "0"),
          "tcz_kzf": int(os.environ.get("MVS_TCZ_KZF", "1") or 0),
          "train_conv": os.environ.get("MVS_TRAIN_CONV", "fp32")}
if _state["train_conv"] not in ("fp32", "tf32x3", "tf32"):
    raise RuntimeError("MVS_TRAIN_CONV must be fp32, tf32x3 or tf32")
if _state["cv_layout"] not in ("cl", "nchw"):
    raise RuntimeError("MVS_CV_LAYOUT must be cl or nchw")
if _state["conv_precision"] not in _VALID:
    raise RuntimeError("MVS_CONV_PRECISION must be one of %s" % (_VALID,))


def conv_precision():
    return _state["conv_precision"]


def set_conv_precision(mode):
    if mode not in _VALID:
        raise ValueError("conv precision must be one of %s, got %r" % (_VALID, mode))
    _state["conv_precision"] = mode


def cv_store():
    return _state["cv_store"]


def set_cv_store(flag):
    _state["cv_store"] = bool(flag)


def cv_layout():
    return _state["cv_layout"]


def set_cv_layout(mode):
    if mode not in ("cl", "nchw"):
        raise ValueError("cv layout must be cl or nchw, got %r" % (mode,))
    _state["cv_layout"] = mode


def vis_fused():
    return _state["vis_fused"]


def set_vis_fused(flag):
    _state["vis_fused"] = bool(flag)


def conv_tma():
    return _state["conv_tma"]


def set_conv_tma(flag):
    _state["conv_tma"] = bool(flag)


def tcz_kzf():
    return _state["tcz_kzf"]


def set_tcz_kzf(level):
    _state["tcz_kzf"] = int(level)


def train_conv():
    return _state["train_conv"]


def set_train_conv(mode):
    if mode not in ("fp32", "tf32x3", "tf32"):
        raise ValueError("train conv mode must be fp32, tf32x3 or tf32, got %r" % (mode,))
    _state["train_conv"] = mode
