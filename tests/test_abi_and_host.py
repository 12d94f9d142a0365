"""CPU-only checks: the C-ABI library loads and exports every declared symbol (no compute call),
the drop-in modules keep the reference's state_dict contract, and the host logic fails loudly
instead of falling back when no GPU is present."""
import ctypes
import json
import os

import pytest
import torch

from mvsformer_b200 import _lib, synthetic as S
from mvsformer_b200 import module as M
from mvsformer_b200.mvsformer_model import CascadeMVS, StageNet
from tests.helpers import GOLDEN, STAGE_ARGS


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    declared = _lib.declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), "libmvs_b200.so does not export %s" % name
    assert set(declared) == set(_lib._SIGNATURES), "python binding table out of sync with include/mvs_b200.h"
    assert lib.mvs_version() == 100
    assert isinstance(lib.mvs_last_error_string(), bytes)


def test_argument_validation_without_gpu():
    """Argument checks run before any CUDA call, so they are testable on a CPU box."""
    lib = _lib.load()
    rc = lib.mvs_relative_projections(None, 1, 5, None, None)
    assert rc == -1 and b"null pointer" in lib.mvs_last_error_string()
    buf = (ctypes.c_float * 64)()
    p = ctypes.cast(buf, ctypes.c_void_p)
    rc = lib.mvs_cost_volume_aggregate(p, 0, 0, p, p, p, p, 1, 5, 30, 8, 4, 8, 8, None)
    assert rc == -1 and b"not divisible" in lib.mvs_last_error_string()
    rc = lib.mvs_conv3d_cl(p, p, None, None, p, 1, 4, 8, 8, 8, 16, 2, 1, 1, 1, 1, None)
    assert rc == -1 and b"kernel size" in lib.mvs_last_error_string()
    rc = lib.mvs_relative_projections(p, 1, 40, p, None)
    assert rc == -1 and b"source views" in lib.mvs_last_error_string()
    # mvs_conv3d_tma_prob (last transposed layer + `prob` conv): null pointers, kernel depth, channel count, TMEM capacity
    rc = lib.mvs_conv3d_tma_prob(p, p, None, None, None, 0.0, p, 1, 4, 8, 8, 16, 3, 1, None)
    assert rc == -1 and b"null pointer" in lib.mvs_last_error_string()
    rc = lib.mvs_conv3d_tma_prob(p, p, None, None, p, 0.0, p, 1, 4, 8, 8, 16, 2, 1, None)
    assert rc == -1 and b"kernel depth" in lib.mvs_last_error_string()
    rc = lib.mvs_conv3d_tma_prob(p, p, None, None, p, 0.0, p, 1, 4, 8, 8, 32, 3, 1, None)
    assert rc < 0 and b"Cin = 16" in lib.mvs_last_error_string()
    rc = lib.mvs_conv3d_tma_prob(p, p, None, None, p, 0.0, p, 1, 16, 8, 8, 16, 3, 1, None)
    assert rc < 0 and b"tensor memory" in lib.mvs_last_error_string()


def test_no_cpu_fallback():
    net = StageNet(dict(STAGE_ARGS), 8, 2).eval()
    feats = torch.zeros(1, 3, 16, 16, 24)
    cams = S.make_cameras(1, 3, 32, 48)["stage3"]
    hyp = torch.ones(1, 8, 16, 24)
    with pytest.raises(RuntimeError, match="CUDA"):
        net(feats, cams, hyp)
    with pytest.raises(RuntimeError, match="CUDA"):
        M.init_inverse_range(S.make_depth_range(1), 32, "cpu", torch.float32, 8, 8)


def _keys_and_shapes(sd):
    return {k: list(v.shape) for k, v in sd.items()}


def test_state_dict_contract_recorded():
    """Key names and shapes of fusions.* as recorded from the reference (tests/golden/state_dict_keys.json,
    written by oracle/make_golden.py); 302 entries, SURVEY.md §5."""
    with open(os.path.join(GOLDEN, "state_dict_keys.json")) as f:
        want = json.load(f)
    args = dict(STAGE_ARGS, ndepths=list(S.NDEPTHS), depth_interals_ratio=list(S.DEPTH_INTERVAL_RATIO), inverse_depth=True)
    got = _keys_and_shapes({k: v for k, v in CascadeMVS(args).state_dict().items()})
    assert got == want
    assert len(got) == 302


def test_costregnet_variants_construct():
    assert M.CostRegNet(8, 8).prob.bias is None and M.CostRegNet(8, 8).prob.kernel_size == (3, 3, 3)
    assert M.CostRegNet3D(8, 8).prob.bias is not None
    assert "conv7.0.weight" in M.CostRegNet3D(8, 8).state_dict()
    assert "conv7.conv.weight" in M.CostRegNet(8, 8).state_dict()
    assert M.CostRegNet2D(8, 8).conv1.conv.kernel_size == (1, 3, 3)


def test_fold_cache_tracks_parameter_versions():
    blk = M.Conv3d(8, 16, padding=1).eval()
    w1, s1 = blk.packed()
    assert blk.packed()[0] is w1                       # cached
    with torch.no_grad():
        blk.bn.running_var.mul_(2.0)
    w2, _ = blk.packed()
    assert w2 is not w1 and not torch.equal(w1, w2)    # rebuilt after an in-place update
    assert tuple(w2.shape) == (3, 3, 3, 8, 16)


def test_algorithmic_bytes_match_survey():
    # cfg 2 = 1009 MB / ref view, 13.27 Mvox
    assert abs(S.cost_volume_algorithmic_bytes(5, 1152, 1536) / 1e6 - 1009) < 2
    assert abs(S.voxels_per_ref_view(1152, 1536) / 1e6 - 13.27) < 0.01


def test_ctypes_signatures_match_header_prototypes():
    """Every prototype of include/mvs_b200.h has a ctypes signature with the same number of parameters and compatible
    kinds (pointer / integer / floating point), and vice versa — drift here would only show up as a crash on the GPU."""
    import ctypes
    import re

    from mvsformer_b200 import _lib

    text = open(_lib.HEADER_PATH).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"^\s*#.*$", "", text, flags=re.M)
    protos = dict(re.findall(r"\b(mvs_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", text))
    assert set(protos) == set(_lib._SIGNATURES), (set(protos) ^ set(_lib._SIGNATURES))
    for name, params in protos.items():
        plist = [p.strip() for p in params.split(",")] if params.strip() not in ("", "void") else []
        _, argtypes = _lib._SIGNATURES[name]
        assert len(plist) == len(argtypes), (name, len(plist), len(argtypes))
        for p, t in zip(plist, argtypes):
            if "*" in p:
                assert t is ctypes.c_void_p or (isinstance(t, type) and issubclass(t, ctypes._Pointer)), (name, p, t)
            elif re.match(r"^(const\s+)?(float|double)\b", p):
                assert t in (ctypes.c_float, ctypes.c_double) and (t is ctypes.c_double) == ("double" in p), (name, p, t)
            else:
                assert t in (ctypes.c_int, ctypes.c_int64, ctypes.c_uint, ctypes.c_longlong), (name, p, t)
                assert (t is ctypes.c_int64) == ("int64_t" in p), (name, p, t)
