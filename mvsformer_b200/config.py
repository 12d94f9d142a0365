"""Run-time knobs of the engine.

``conv_precision`` selects the arithmetic of the 3D-CNN convolutions (mvs_conv3d_* / mvs_deconv3d_*):

* ``"tf32x3"`` (default) — tcgen05 tensor cores, each operand split into two TF32 terms and three
  MMAs per product (error ~2^-21: fp32-grade; keeps the fp32 parity gates of the tests).
* ``"tf32"``   — tcgen05 tensor cores, operands rounded to TF32 (what cuDNN does by default for
  the reference on a GPU, ``torch.backends.cudnn.allow_tf32 = True``).
* ``"fp32"``   — FP32 CUDA-core kernels (exact fp32 FMA chains).

Set with ``set_conv_precision()`` or the environment variable ``MVS_CONV_PRECISION``.

``cv_layout`` (``MVS_CV_LAYOUT`` = ``cl`` (default) | ``nchw``) — which cost-volume kernels sample the features: the
channels-last kernels (csrc/cost_volume_cl.cu: one LDS.128 per 4-channel tap, conflict-free lane order, next item's TMA
tile in flight, one sampling pass at stages 1-3; features are re-laid out once per call by mvs_features_to_cl) or the
generic any-shape kernels (csrc/cost_volume.cu) that read the reference's NCHW layout directly — they are also the
fall-back for shapes the channels-last kernels are not built for, and the second implementation the parity tests
cross-check against.

``vis_fused`` (``MVS_VIS_FUSED``, default 1) — the visibility net as one persistent kernel (csrc/vis_fused.cu, TF32 mode):
only the entropy map and the weight map touch HBM; ``0`` = the FP32 CUDA-core kernel of csrc/vis_net.cu (also what the
``tf32x3`` and ``fp32`` modes use).

``conv_tma`` (``MVS_CONV_TMA``, default 1; ``0`` = the generic tcgen05 kernels of csrc/conv3d_tc.cu only) — depth-unstrided
3x3x3 layers (CostRegNet3D) through the persistent, warp-specialised, TMA-fed tcgen05 kernels of csrc/conv3d_tma.cu
(TF32 mode only).

``prob_fused`` (``MVS_PROB_FUSED``, default 1) — eval, TF32 mode: CostRegNet3D's last transposed layer (16 -> 8) applies the
1x1x1 ``prob`` conv in its epilogue (mvs_conv3d_tma_prob), so the regulariser's 8-channel output tensor never reaches HBM;
bit-identical to the two-kernel route (``0``).

``train_conv`` (``MVS_TRAIN_CONV`` = ``fp32`` (default) | ``tf32x3`` | ``tf32``) — arithmetic of the training path's
forward and data-gradient convolutions: the FP32 CUDA-core kernels, or the tcgen05 kernels of the inference path
(``tf32x3`` keeps fp32-grade accuracy; the reference itself trains the regulariser under fp16 autocast).  Weight
gradients stay FP32.  Measured on B200 at the cfg-5 shape (profiles/r02_train_scale.json): fp32 34.96 ms per step, tf32
43.8 ms (the per-layer operand repacking costs more than the tensor cores save at these sizes), so fp32 stays the default.
"""
import os

_VALID = ("tf32x3", "tf32", "fp32")
_state = {"conv_precision": os.environ.get("MVS_CONV_PRECISION", "tf32x3"),
          "cv_layout": os.environ.get("MVS_CV_LAYOUT", "cl"),
          "vis_fused": os.environ.get("MVS_VIS_FUSED", "1") not in ("", "0"),
          "conv_tma": os.environ.get("MVS_CONV_TMA", "1") not in ("", "0"),
          "prob_fused": os.environ.get("MVS_PROB_FUSED", "1") not in ("", "0"),
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


def prob_fused():
    return _state["prob_fused"]


def set_prob_fused(flag):
    _state["prob_fused"] = bool(flag)


def train_conv():
    return _state["train_conv"]


def set_train_conv(mode):
    if mode not in ("fp32", "tf32x3", "tf32"):
        raise ValueError("train conv mode must be fp32, tf32x3 or tf32, got %r" % (mode,))
    _state["train_conv"] = mode
