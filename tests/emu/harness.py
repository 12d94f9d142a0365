"""Points the package's ctypes layer at the CPU emulation library (tests only).

``install(setattr_fn)`` takes pytest's ``monkeypatch.setattr`` (or plain ``setattr`` inside a spawned
worker process) and swaps: the library handle, the CUDA-tensor checks (host tensors are what the
emulation works on), the stream getter, and the two entry points of the training step that are not
part of the emulated set (relative projections and the head, restated from the oracle)."""
import ctypes

import torch

from mvsformer_b200 import _lib, engine
from oracle import mvs_oracle as O
from tests.emu import build_emu

EMU_SYMBOLS = [
    "mvs_last_error_string", "mvs_launch_count", "mvs_conv3d_cl", "mvs_deconv3d_cl", "mvs_group_corr_fwd",
    "mvs_group_corr_bwd", "mvs_corr_entropy", "mvs_aggregate_fwd", "mvs_aggregate_bwd", "mvs_bn_stats", "mvs_bn_collapse", "mvs_bn_finalize",
    "mvs_bn_act_fwd", "mvs_bn_act_bwd_reduce", "mvs_bn_act_bwd_apply", "mvs_conv_wgrad_cl", "mvs_thin_conv_cl",
    "mvs_sigmoid_bwd", "mvs_softmax_bwd", "mvs_proj_mask", "mvs_epipole_aggregate_fwd", "mvs_epipole_aggregate_bwd", "mvs_homo_warp", "mvs_homo_warp_bwd", "mvs_homo_warp_bwd_grid", "mvs_depth_regression_bwd", "mvs_mixup_head",
    "mvs_fusion_reproject", "mvs_fusion_reproject_dynamic", "mvs_fusion_filter_dynamic", "mvs_fusion_filter", "mvs_fusion_points", "mvs_fusion_prob_filter"]


def _host_only(*tensors):
    for t in tensors:
        if t is None:
            continue
        assert not t.is_cuda and t.dtype in (torch.float32, torch.float64) and t.is_contiguous(), (t.dtype, t.stride())


def _relproj(proj):
    b, v = proj.shape[:2]
    ref = O.compose_projection(proj[:, 0].double())
    rows = []
    for i in range(1, v):
        rot, trans = O.relative_projection(O.compose_projection(proj[:, i].double()), ref)
        rows.append(torch.cat([rot, trans.unsqueeze(-1)], dim=-1).reshape(b, 12))
    return torch.stack(rows, dim=1).float().contiguous()


def _head(pre, depth_values, tmp, training, want_prob=True):
    prob, depth, conf = O.regression_head(pre.detach(), depth_values, tmp, training)
    return prob.contiguous(), depth.contiguous(), conf.contiguous()


def install(setattr_fn):
    lib = ctypes.CDLL(build_emu.build())
    for name in EMU_SYMBOLS:
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = _lib._SIGNATURES[name]
    setattr_fn(_lib, "load", lambda: lib)
    setattr_fn(_lib, "require_cuda", _host_only)
    setattr_fn(_lib, "stream", lambda: None)
    setattr_fn(engine, "require_cuda", _host_only)
    setattr_fn(engine, "stream", lambda: None)
    setattr_fn(engine, "relative_projections", _relproj)
    setattr_fn(engine, "regression_head", _head)
    setattr_fn(engine, "init_range", lambda cur, nd, h, w, inverse: (O.init_inverse_range if inverse else O.init_range)(
        cur.float(), nd, h, w).contiguous())
    setattr_fn(engine, "schedule_inverse_range", lambda depth, hypo, nd, split, h, w: O.schedule_inverse_range(
        depth.float(), hypo.float(), nd, split, h, w).contiguous())
    setattr_fn(engine, "depth_regression", lambda p, dv: O.depth_regression(p.detach(), dv).contiguous())
    setattr_fn(engine, "conf_regression", lambda p, n: O.conf_regression(p.detach(), n).contiguous())
    return lib
