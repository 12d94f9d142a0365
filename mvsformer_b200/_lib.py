"""ctypes binding of libmvs_b200.so (C ABI declared in include/mvs_b200.h).

The library is the product: if it is missing, or a tensor is not a CUDA fp32 tensor, the
wrappers raise — there is no CPU or PyTorch fallback anywhere in this package.
"""
import ctypes
import os
import re

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# MVS_LIB_PATH: load another build of the SAME library (A/B of kernel generations, scripts/ab_libs.sh); a build that lacks a
# declared entry point still loads — calling the missing entry point raises
LIB_PATH = os.environ.get("MVS_LIB_PATH") or os.path.join(_HERE, "lib", "libmvs_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "mvs_b200.h")

_lib = None

c_f = ctypes.c_void_p      # device / host float pointers travel as void*
c_i = ctypes.c_int
c_l = ctypes.c_int64
c_fl = ctypes.c_float
c_d = ctypes.c_double

_SIGNATURES = {
    "mvs_version": (c_i, []),
    "mvs_last_error_string": (ctypes.c_char_p, []),
    "mvs_launch_count": (ctypes.c_longlong, []),
    "mvs_device_info": (c_i, [ctypes.POINTER(c_i)] * 3),
    "mvs_relative_projections": (c_i, [c_f, c_i, c_i, c_f, c_f]),
    "mvs_relative_projection_pair": (c_i, [c_f, c_f, c_i, c_f, c_f]),
    "mvs_homo_warp": (c_i, [c_f, c_f, c_f, c_i, c_f, c_f, c_i, c_i, c_i, c_i, c_i, c_f]),
    "mvs_cost_volume_entropy": (c_i, [c_f, c_l, c_l, c_f, c_f, c_f, c_f] + [c_i] * 7 + [c_f]),
    "mvs_cost_volume_aggregate": (c_i, [c_f, c_l, c_l, c_f, c_f, c_f, c_f] + [c_i] * 7 + [c_f]),
    "mvs_cost_volume_aggregate_tf32": (c_i, [c_f, c_l, c_l, c_f, c_f, c_f, c_f] + [c_i] * 7 + [c_f]),
    "mvs_features_to_cl": (c_i, [c_f, c_f, c_f, c_f, c_f, c_i, c_f]),
    "mvs_cost_volume_cl_entropy": (c_i, [c_f, c_i, c_f] + [c_f] * 5 + [c_i] * 7 + [c_f]),
    "mvs_cost_volume_cl_aggregate": (c_i, [c_f, c_i, c_f] + [c_f] * 4 + [c_i] * 8 + [c_f]),
    "mvs_corr_aggregate": (c_i, [c_f, c_f, c_f] + [c_i] * 6 + [c_f]),
    "mvs_argmax_gather": (c_i, [c_f, c_f, c_f, c_i, c_i, c_i, c_i, c_f]),
    "mvs_vis_weight": (c_i, [c_f, c_f, c_f, c_i, c_i, c_i, c_f]),
    "mvs_vis_fused": (c_i, [c_f] * 5 + [c_i] * 3 + [c_f]),
    "mvs_conv3d_cl": (c_i, [c_f] * 5 + [c_i] * 11 + [c_f]),
    "mvs_deconv3d_cl": (c_i, [c_f] * 5 + [c_i] * 9 + [c_f]),
    "mvs_conv3d_tc": (c_i, [c_f] * 6 + [c_i] * 11 + [c_f]),
    "mvs_deconv3d_tc": (c_i, [c_f] * 6 + [c_i] * 10 + [c_f]),
    "mvs_conv3d_tma": (c_i, [c_f] * 5 + [c_i] * 10 + [c_f]),
    "mvs_conv3d_tma_prob": (c_i, [c_f] * 5 + [c_fl, c_f] + [c_i] * 7 + [c_f]),
    "mvs_ncdhw_to_cl_tf32": (c_i, [c_f, c_f] + [c_i] * 5 + [c_f]),
    "mvs_tc_probe": (c_i, [c_f, c_i, c_f, c_i] + [ctypes.c_uint] * 4 + [c_i, c_i, ctypes.c_uint, ctypes.c_uint, c_f, c_f]),
    "mvs_tc_probe_ts": (c_i, [c_f, c_i, c_f, c_i] + [ctypes.c_uint] * 4 + [c_i, c_i, ctypes.c_uint, ctypes.c_uint, ctypes.c_uint, c_f, c_f]),
    "mvs_ncdhw_to_cl": (c_i, [c_f, c_f] + [c_i] * 5 + [c_f]),
    "mvs_cl_to_ncdhw": (c_i, [c_f, c_f] + [c_i] * 5 + [c_f]),
    "mvs_prob_conv_cl": (c_i, [c_f, c_f, c_f, c_f] + [c_i] * 6 + [c_f]),
    "mvs_regression_head": (c_i, [c_f, c_f, c_fl, c_i, c_f, c_f, c_f] + [c_i] * 4 + [c_f]),
    "mvs_depth_regression": (c_i, [c_f, c_f, c_i, c_f] + [c_i] * 4 + [c_f]),
    "mvs_conf_regression": (c_i, [c_f, c_i, c_f] + [c_i] * 4 + [c_f]),
    "mvs_init_inverse_range": (c_i, [c_f, c_i, c_f] + [c_i] * 4 + [c_f]),
    "mvs_init_range": (c_i, [c_f, c_i, c_f] + [c_i] * 4 + [c_f]),
    "mvs_schedule_inverse_range": (c_i, [c_f, c_f, c_i, c_fl, c_f] + [c_i] * 4 + [c_f]),
    "mvs_schedule_range": (c_i, [c_f, c_f, c_f] + [c_i] * 4 + [c_f]),
    "mvs_confidence_accumulate": (c_i, [c_f, c_i, c_i, c_f, c_i, c_i, c_i, c_fl, c_f]),
    "mvs_confidence_upsample_accumulate": (c_i, [c_f, c_i, c_i, c_f, c_f, c_i, c_i, c_i, c_fl, c_f]),
    # training path (train.cu)
    "mvs_group_corr_fwd": (c_i, [c_f, c_l, c_l, c_f, c_f, c_f] + [c_i] * 7 + [c_f]),
    "mvs_group_corr_bwd": (c_i, [c_f, c_l, c_l, c_f, c_f, c_f, c_f] + [c_i] * 7 + [c_f]),
    "mvs_corr_entropy": (c_i, [c_f, c_f] + [c_i] * 5 + [c_f]),
    "mvs_aggregate_fwd": (c_i, [c_f, c_f, c_f] + [c_i] * 6 + [c_f]),
    "mvs_aggregate_bwd": (c_i, [c_f] * 5 + [c_i] * 6 + [c_f]),
    "mvs_bn_stats": (c_i, [c_f, c_f, c_l, c_i, c_f]),
    "mvs_bn_collapse": (c_i, [c_f, c_i, c_f]),
    "mvs_bn_finalize": (c_i, [c_f, c_i, c_d, c_fl, c_fl, c_f, c_f, c_f, c_i, c_f]),
    "mvs_bn_act_fwd": (c_i, [c_f] * 6 + [c_l, c_i, c_i, c_f]),
    "mvs_bn_act_bwd_reduce": (c_i, [c_f] * 6 + [c_l, c_i, c_i, c_f]),
    "mvs_bn_act_bwd_apply": (c_i, [c_f] * 6 + [c_d, c_f, c_l, c_i, c_i, c_f]),
    "mvs_conv_wgrad_cl": (c_i, [c_f, c_f, c_f] + [c_i] * 14 + [c_f]),
    "mvs_thin_conv_cl": (c_i, [c_f] * 4 + [c_i] * 9 + [c_f]),
    "mvs_sigmoid_bwd": (c_i, [c_f, c_f, c_f, c_l, c_f]),
    "mvs_homo_warp_bwd": (c_i, [c_f, c_f, c_f, c_i, c_f] + [c_i] * 5 + [c_f]),
    "mvs_proj_mask": (c_i, [c_f, c_f, c_f] + [c_i] * 5 + [c_f]),
    "mvs_epipole_aggregate_fwd": (c_i, [c_f, c_f, c_fl, c_fl, c_f, c_f, c_f] + [c_i] * 6 + [c_f]),
    "mvs_epipole_aggregate_bwd": (c_i, [c_f] * 6 + [c_fl, c_fl, c_f, c_f] + [c_i] * 6 + [c_f]),
    "mvs_homo_warp_bwd_grid": (c_i, [c_f] * 4 + [c_i] + [c_f, c_f] + [c_i] * 5 + [c_f]),
    "mvs_depth_regression_bwd": (c_i, [c_f, c_f, c_i, c_f] + [c_i] * 4 + [c_f]),
    "mvs_mixup_head": (c_i, [c_f] * 4 + [c_i] * 4 + [c_f]),
    "mvs_softmax_bwd": (c_i, [c_f, c_f, c_f] + [c_i] * 4 + [c_f]),
    # depth-map fusion (fusion.cu)
    "mvs_fusion_reproject": (c_i, [c_f] * 5 + [c_i] * 4 + [c_f]),
    "mvs_fusion_filter": (c_i, [c_f] * 3 + [c_fl] * 3 + [c_f] * 3 + [c_i] * 4 + [c_f]),
    "mvs_fusion_reproject_dynamic": (c_i, [c_f] * 4 + [c_i] * 4 + [c_f]),
    "mvs_fusion_filter_dynamic": (c_i, [c_f, c_f, c_fl, c_fl] + [c_f] * 4 + [c_i] * 4 + [c_f]),
    "mvs_fusion_points": (c_i, [c_f] * 3 + [c_i] * 3 + [c_f]),
    "mvs_fusion_prob_filter": (c_i, [c_f, c_f, c_i, c_f] + [c_i] * 4 + [c_f]),
}


def declared_symbols():
    """Every function name declared in include/mvs_b200.h (used by the ABI test)."""
    text = open(HEADER_PATH).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mvs_[a-z0-9_]+)\s*\(", text)))


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libmvs_b200.so is not built (%s). Run `python -m mvsformer_b200.build` "
            "(or __graft_entry__.build()). There is no fallback path." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        if not hasattr(lib, name) and os.environ.get("MVS_LIB_PATH"):
            continue
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().mvs_last_error_string().decode("utf-8", "replace")
        raise RuntimeError("%s failed (%d): %s" % (what or "libmvs_b200", rc, msg))


def require_cuda(*tensors):
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("mvsformer_b200 runs on CUDA tensors only (got a %s tensor); there is no CPU path"
                               % t.device.type)
        check_device(t)
        if t.dtype != torch.float32:
            raise RuntimeError("expected a float32 tensor, got %s" % t.dtype)
        if not t.is_contiguous():
            raise RuntimeError("expected a contiguous tensor (shape %s, strides %s)" % (tuple(t.shape), t.stride()))


def check_device(t):
    """Kernels are enqueued on the CURRENT device's current stream (the C side never calls cudaSetDevice): a tensor
    on another GPU would be launched against the wrong context, so refuse it."""
    if t.device.index != torch.cuda.current_device():
        raise RuntimeError("tensor on %s but the current CUDA device is cuda:%d; call inside `with torch.cuda.device(%d)`"
                           % (t.device, torch.cuda.current_device(), t.device.index))


def ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
