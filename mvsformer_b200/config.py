"""Run-time knobs of the engine.

``conv_precision`` selects the arithmetic of the 3D-CNN convolutions (mvs_conv3d_* / mvs_deconv3d_*):

* ``"tf32x3"`` (default) — tcgen05 tensor cores, each operand split into two TF32 terms and three
  MMAs per product (error ~2^-21: fp32-grade; keeps the fp32 parity gates of the tests).
* ``"tf32"``   — tcgen05 tensor cores, operands rounded to TF32 (what cuDNN does by default for
  the reference on a GPU, ``torch.backends.cudnn.allow_tf32 = True``).
* ``"fp32"``   — FP32 CUDA-core kernels (exact fp32 FMA chains).

Set with ``set_conv_precision()`` or the environment variable ``MVS_CONV_PRECISION``.
"""
import os

_VALID = ("tf32x3", "tf32", "fp32")
_state = {"conv_precision": os.environ.get("MVS_CONV_PRECISION", "tf32x3")}
if _state["conv_precision"] not in _VALID:
    raise RuntimeError("MVS_CONV_PRECISION must be one of %s" % (_VALID,))


def conv_precision():
    return _state["conv_precision"]


def set_conv_precision(mode):
    if mode not in _VALID:
        raise ValueError("conv precision must be one of %s, got %r" % (_VALID, mode))
    _state["conv_precision"] = mode
