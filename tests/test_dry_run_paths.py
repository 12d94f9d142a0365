"""Dry run of the host code of every forward path on the CPU: the library handle is replaced by a recorder that checks each
call against the declared ctypes signature (argument count and convertibility) and launches nothing.  Outputs are
uninitialised memory, so nothing numerical is asserted — the test exists to catch host-side mistakes (wrong argument
lists, shape logic, missing attributes) in code paths that only run on a GPU: inference in all three precision modes,
with every knob of config.py (MVS_CV_LAYOUT, MVS_CONV_TMA, MVS_VIS_FUSED, MVS_TRAIN_CONV)."""
import ctypes

import pytest
import torch

from mvsformer_b200 import _lib, config, engine
from mvsformer_b200 import synthetic as S
from mvsformer_b200.mvsformer_model import CascadeMVS
from tests.helpers import CASCADE_ARGS


class _Recorder:
    def __init__(self):
        self.calls = []

    def __getattr__(self, name):
        if name not in _lib._SIGNATURES:
            raise AttributeError("%s is not declared in _lib._SIGNATURES" % name)
        restype, argtypes = _lib._SIGNATURES[name]

        def fn(*args):
            assert len(args) == len(argtypes), "%s: %d arguments, signature has %d" % (name, len(args), len(argtypes))
            for a, t in zip(args, argtypes):
                if t is ctypes.c_void_p:
                    assert a is None or isinstance(a, (int, ctypes.c_void_p)) or hasattr(a, "_as_parameter_") \
                        or isinstance(a, ctypes._SimpleCData) or isinstance(a, ctypes.Array), (name, type(a))
                else:
                    t(a)                                         # raises TypeError when not convertible
            self.calls.append(name)
            return b"dry run" if restype is ctypes.c_char_p else 0
        return fn


@pytest.fixture()
def dry(monkeypatch):
    rec = _Recorder()
    monkeypatch.setattr(_lib, "load", lambda: rec)
    monkeypatch.setattr(_lib, "require_cuda", lambda *t: None)
    monkeypatch.setattr(_lib, "stream", lambda: None)
    monkeypatch.setattr(engine, "require_cuda", lambda *t: None)
    monkeypatch.setattr(engine, "stream", lambda: None)
    monkeypatch.setattr(engine, "require_cuda_device", lambda t: None)
    return rec


def _cascade_inputs(height=128, width=256, views=3):
    feats = S.make_features(1, views, height, width, seed=3)
    cams = S.make_cameras(1, views, height, width)
    return feats, cams, S.make_depth_range(1)


@pytest.mark.parametrize("mode", ["tf32", "tf32x3", "fp32"])
@pytest.mark.parametrize("layout,tma,fused", [("cl", True, True), ("cl", False, False), ("nchw", True, True), ("nchw", False, False)])
def test_inference_host_path(dry, mode, layout, tma, fused):
    feats, cams, dv = _cascade_inputs()
    net = CascadeMVS(dict(CASCADE_ARGS)).eval()
    old = config.conv_precision()
    config.set_conv_precision(mode)
    config.set_cv_layout(layout)
    config.set_conv_tma(tma)
    config.set_vis_fused(fused)
    try:
        with torch.no_grad():
            out = net(feats, cams, dv, tmp=list(S.EVAL_TMP))
    finally:
        config.set_conv_precision(old)
        config.set_cv_layout("cl")
        config.set_conv_tma(True)
        config.set_vis_fused(True)
    assert out["refined_depth"].shape == (1, 128, 256) and out["photometric_confidence"].shape == (1, 128, 256)
    assert set(out["stage1"]) >= {"depth", "prob_volume", "photometric_confidence", "depth_values", "prob_volume_pre", "sim_depth"}
    called = set(dry.calls)
    if layout == "cl":                   # default: channels-last kernels, one sampling pass at stages 1-3
        assert {"mvs_features_to_cl", "mvs_cost_volume_cl_entropy", "mvs_cost_volume_cl_aggregate", "mvs_corr_aggregate"} <= called
        assert dry.calls.count("mvs_cost_volume_cl_entropy") == 4 and dry.calls.count("mvs_cost_volume_cl_aggregate") == 1
        assert not called & {"mvs_cost_volume_entropy", "mvs_cost_volume_aggregate", "mvs_cost_volume_aggregate_tf32"}
    else:                                # generic kernels on the reference's NCHW layout, two sampling passes
        assert dry.calls.count("mvs_cost_volume_entropy") == 4 and "mvs_features_to_cl" not in called
        assert dry.calls.count("mvs_cost_volume_aggregate" + ("_tf32" if mode == "tf32" else "")) == 4
        assert dry.calls.count("mvs_argmax_gather") == 4
    if mode == "tf32" and tma:           # depth-unstrided layers on the persistent TMA kernels, the rest on the generic tcgen05 ones
        # stages 3-4: the last transposed layer takes the 1x1x1 `prob` conv into its epilogue (config.prob_fused)
        assert dry.calls.count("mvs_conv3d_tma") == 22 and dry.calls.count("mvs_conv3d_tma_prob") == 2
        assert dry.calls.count("mvs_prob_conv_cl") == 2
        assert dry.calls.count("mvs_conv3d_tc") + dry.calls.count("mvs_deconv3d_tc") == 12
    else:
        assert "mvs_conv3d_tma" not in called and "mvs_conv3d_tma_prob" not in called
        assert dry.calls.count("mvs_prob_conv_cl") == 4
    if mode == "tf32" and fused:
        assert dry.calls.count("mvs_vis_fused") == 4 and "mvs_vis_weight" not in called
    else:
        assert dry.calls.count("mvs_vis_weight") == 4 and "mvs_vis_fused" not in called
    if mode == "fp32":
        assert "mvs_conv3d_cl" in called and not any("_tc" in c or "_tma" in c for c in called)
    elif not (mode == "tf32" and tma):
        assert dry.calls.count("mvs_conv3d_tc") + dry.calls.count("mvs_deconv3d_tc") == 36


def test_prob_fused_knob_off_keeps_the_two_kernel_route(dry):
    feats, cams, dv = _cascade_inputs()
    net = CascadeMVS(dict(CASCADE_ARGS)).eval()
    old = config.conv_precision()
    config.set_conv_precision("tf32")
    config.set_prob_fused(False)
    try:
        with torch.no_grad():
            net(feats, cams, dv, tmp=list(S.EVAL_TMP))
    finally:
        config.set_conv_precision(old)
        config.set_prob_fused(True)
    assert dry.calls.count("mvs_conv3d_tma") == 24 and "mvs_conv3d_tma_prob" not in dry.calls
    assert dry.calls.count("mvs_prob_conv_cl") == 4


@pytest.mark.parametrize("train_conv", ["fp32", "tf32x3", "tf32"])
def test_training_host_path(dry, train_conv):
    feats, cams, dv = _cascade_inputs(64, 128)
    feats = {k: v.requires_grad_(True) for k, v in feats.items()}
    net = CascadeMVS(dict(CASCADE_ARGS)).train()
    config.set_train_conv(train_conv)
    try:
        out = net(feats, cams, dv)
        loss = sum(out["stage%d" % (s + 1)]["prob_volume_pre"].float().mean() for s in range(4))
        loss.backward()
    finally:
        config.set_train_conv("fp32")
    called = set(dry.calls)
    assert {"mvs_group_corr_fwd", "mvs_group_corr_bwd", "mvs_aggregate_bwd", "mvs_conv_wgrad_cl", "mvs_bn_act_bwd_apply"} <= called
    if train_conv != "fp32":
        assert "mvs_conv3d_tc" in called and "mvs_deconv3d_tc" in called
    assert all(p.grad is not None for p in net.parameters())


@pytest.mark.parametrize("kind", ["epipole", "epipoleV2"])
@pytest.mark.parametrize("training", [False, True])
def test_epipole_host_path(dry, kind, training):
    from mvsformer_b200.mvsformer_model import StageNet
    from tests.helpers import STAGE_ARGS

    feats = S.make_features(1, 3, 64, 96, seed=1, stages=(2,))["stage3"]
    cams = S.make_cameras(1, 3, 64, 96)["stage3"]
    hyp = S.narrow_hypotheses(2, 64, 96, 1)
    net = StageNet(dict(STAGE_ARGS, fusion_type=kind), 8, 2)
    net = net.train() if training else net.eval()
    old = config.conv_precision()
    config.set_conv_precision("tf32")
    try:
        out = net(feats, cams, hyp, tmp=[5.0, 5.0, 5.0, 1.0])
    finally:
        config.set_conv_precision(old)
    assert out["prob_volume_pre"].shape == (1, 8, 32, 48) and ("sim_depth" in out) == (not training)
    assert "mvs_epipole_aggregate_fwd" in dry.calls and (("mvs_proj_mask" in dry.calls) == (kind == "epipoleV2"))


def test_other_module_entry_points_host_path(dry):
    from mvsformer_b200 import fusion as Fu
    from mvsformer_b200 import module as M
    from mvsformer_b200 import warping as Wp

    x = torch.zeros(1, 8, 8, 16, 24)
    for cls in (M.CostRegNet, M.CostRegNet3D, M.CostRegNet2D):
        assert cls(8, 8).eval()(x).shape == (1, 1, 8, 16, 24)
    p, dv = torch.zeros(1, 8, 4, 6), torch.zeros(1, 8)
    assert M.depth_regression(p, dv).shape == (1, 4, 6) and M.conf_regression(p, 2).shape == (1, 4, 6)
    # the broadcast forms the reference accepts (models/module.py:597-603) are expanded, anything else is refused
    p2 = torch.zeros(2, 8, 4, 6)
    for form in (torch.zeros(8), torch.zeros(1, 8), torch.zeros(2, 8, 1, 1), torch.zeros(2, 8, 4, 6)):
        assert M.depth_regression(p2, form).shape == (2, 4, 6)
    for bad in (torch.zeros(2, 7), torch.zeros(3, 8), torch.zeros(2, 8, 4, 5)):
        with pytest.raises(RuntimeError):
            M.depth_regression(p2, bad)
    assert M.schedule_range(torch.ones(1, 8, 12), 16, torch.ones(1), 16, 24).shape == (1, 16, 16, 24)
    with pytest.raises(RuntimeError):
        M.schedule_range(torch.ones(1, 16, 24), 16, torch.ones(1), 16, 24)          # same-size map: not [B,H/2,W/2]
    assert M.init_inverse_range(torch.ones(1, 192), 32, None, None, 4, 6).shape == (1, 32, 4, 6)
    eye = torch.eye(4).unsqueeze(0)
    warped, mask = Wp.homo_warping_3D_with_mask(torch.zeros(1, 8, 4, 6), eye, eye, dv)
    assert warped.shape == (1, 8, 8, 4, 6) and mask.shape == (1, 8, 4, 6)
    c = S.make_fusion_case(3, 8, 12)
    out = Fu.filter_view(c["ref_depth"], c["src_depths"], c["ref_cam"], c["src_cams"], 1.0, 0.01, 2,
                         ref_conf=c["ref_conf"], prob_thresh=[0.1, 0.2, 0.3])
    assert out["points"].shape == (1, 3, 8, 12)
    assert Fu.dynamic_filter_view(c["ref_depth"], c["src_depths"], c["ref_cam"], c["src_cams"])["depth_ave"].shape == (1, 1, 8, 12)


def test_bench_kernel_profiler_wraps_existing_entry_points():
    """bench.py's per-kernel attribution wraps engine functions by name: every name must exist."""
    import bench

    prof = bench.KernelProfiler()
    before = {k: getattr(engine, k) for k in dir(engine) if callable(getattr(engine, k))}
    prof.install(engine)
    assert {"features_to_cl", "cost_volume_cl_entropy", "cost_volume_cl_aggregate", "corr_aggregate", "conv3d_tma", "vis_fused",
            "conv3d_tc", "deconv3d_tc", "cost_volume_entropy", "cost_volume_aggregate", "vis_weight"} <= set(prof._saved)
    prof.uninstall(engine)
    assert all(getattr(engine, k) is v for k, v in before.items())


@pytest.mark.parametrize("layout", ["cl", "nchw"])
def test_bench_kernel_profiler_cost_functions(dry, monkeypatch, layout):
    """The profiler's algorithmic-byte / flop lambdas must accept the argument lists the engine wrappers are called
    with, for the shipped path and for the generic kernels."""
    import bench

    class _Event:
        def __init__(self, enable_timing=False):
            pass

        def record(self):
            pass

        def elapsed_time(self, other):
            return 1.0

    monkeypatch.setattr(torch.cuda, "Event", _Event)
    feats, cams, dv = _cascade_inputs()
    net = CascadeMVS(dict(CASCADE_ARGS)).eval()
    old = config.conv_precision()
    config.set_conv_precision("tf32")
    config.set_cv_layout(layout)
    prof = bench.KernelProfiler()
    prof.install(engine)
    try:
        with torch.no_grad():
            net(feats, cams, dv, tmp=list(S.EVAL_TMP))
    finally:
        prof.uninstall(engine)
        config.set_conv_precision(old)
        config.set_cv_layout("cl")
    summ = prof.summary(1)
    assert all(v["alg_bytes_per_step"] > 0 for v in summ.values())
    if layout == "cl":       # the channels-last kernels (one sampling pass at stages 1-3 + streaming aggregation)
        assert {"cv_layout(nchw->channels-last)", "cv_cl_passA+store", "cv_cl_passA(stage4)", "cv_cl_passB(stage4)",
                "cv_corr_aggregate(stream)"} <= set(summ)
    else:
        assert {"cv_entropy(passA)", "cv_aggregate(passB)"} <= set(summ)
    assert "conv3d_tma" in summ and "vis_net(fused)" in summ


def test_prob_fusion_is_limited_to_the_layers_the_fused_kernel_implements():
    """module._RegNetBase._prob_fusable: only CostRegNet3D's 3x3x3 16 -> 8 transposed last layer + 1x1x1 prob conv, in eval,
    TF32 mode, with the persistent TMA kernels and the knob on."""
    from mvsformer_b200 import module as M
    y = torch.zeros(1, 4, 16, 16, 16)
    net3d, net, net2d = M.CostRegNet3D(8, 8).eval(), M.CostRegNet(8, 8).eval(), M.CostRegNet2D(8, 8).eval()
    old = config.conv_precision()
    try:
        config.set_conv_precision("tf32")
        assert net3d._prob_fusable(y)
        assert not net._prob_fusable(y)                    # 3x3x3 prob conv, depth-strided transposed layer
        assert not net2d._prob_fusable(y)                  # (1,3,3) transposed layer: not the kernel _run_deconv would pick
        assert not net3d.train()._prob_fusable(y)
        net3d.eval()
        assert not net3d._prob_fusable(torch.zeros(1, 16, 16, 16, 16))     # 4 classes x 16 slices x 16 columns > tensor memory
        config.set_prob_fused(False)
        assert not net3d._prob_fusable(y)
        config.set_prob_fused(True)
        config.set_conv_tma(False)
        assert not net3d._prob_fusable(y)
        config.set_conv_tma(True)
        for mode in ("tf32x3", "fp32"):
            config.set_conv_precision(mode)
            assert not net3d._prob_fusable(y)
    finally:
        config.set_conv_precision(old)
        config.set_prob_fused(True)
        config.set_conv_tma(True)
