"""GPU parity: every entry point of libmvs_b200.so (called through the C ABI via the Python mirror)
against the CPU oracle on the same seeded inputs, and against the golden vectors produced by the
unmodified reference (tests/golden, oracle/make_golden.py).

Tolerances (fp32 path, stated per check):
* index / mask outputs: agreement rate (pixels within 1 ulp of a border may flip)
* sampled / correlated quantities: relative L1 <= 2e-5 on smooth features (white-noise features
  amplify the 5e-5 px coordinate rounding of the reference's own normalise/un-normalise round trip)
* depth maps: relative L1 <= 1e-3 is the north-star bar; we assert 1e-5 here.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from mvsformer_b200 import engine, module as M, synthetic as S, warping as Wp
from mvsformer_b200.mvsformer_model import CascadeMVS, StageNet
from oracle import mvs_oracle as O
from tests.helpers import STAGE_ARGS, checksum, load_golden, max_abs, rel_l1

pytestmark = pytest.mark.gpu
DEV = "cuda"


def cu(t):
    return t.to(DEV)


# ------------------------------------------------------------------------------------------------
def test_library_reports_blackwell():
    import ctypes
    from mvsformer_b200 import _lib
    lib = _lib.load()
    sm, maj, mnr = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    _lib.check(lib.mvs_device_info(ctypes.byref(sm), ctypes.byref(maj), ctypes.byref(mnr)))
    assert maj.value == 10 and sm.value >= 100


def test_relative_projections():
    cams = S.make_cameras(2, 5, 128, 192)["stage3"]
    got = engine.relative_projections(cu(cams)).cpu()
    ref_p = O.compose_projection(cams[:, 0].double())
    for v in range(1, 5):
        rot, trans = O.relative_projection(O.compose_projection(cams[:, v].double()), ref_p)
        want = torch.cat([rot, trans.unsqueeze(-1)], dim=-1).reshape(2, 12)
        assert rel_l1(got[:, v - 1], want) < 1e-6


def test_warp_vs_golden_and_oracle():
    g = load_golden("warp.npz")
    feats = S.make_features(2, 3, 24, 40, seed=5, stages=(3,), feat_chs=(0, 0, 0, 8))["stage4"]
    assert checksum(feats) == pytest.approx(float(g["feat_checksum"]), rel=1e-12)
    cams = S.make_cameras(2, 3, 24, 40)["stage4"].clone()
    cams[:, 2, 0, 0, 3] += 150.0
    ref_p = O.compose_projection(cams[:, 0])
    dv_map = S.make_depth_range(2)[:, ::32][:, :6]
    dv_px = dv_map.view(2, 6, 1, 1) * (1.0 + 0.1 * torch.rand(2, 6, 24, 40, generator=S._gen(3)))
    for v in (1, 2):
        src_p = O.compose_projection(cams[:, v])
        for tag, dv in (("bd", dv_map), ("px", dv_px)):
            warped, mask = Wp.homo_warping_3D_with_mask(cu(feats[:, v]), cu(src_p), cu(ref_p), cu(dv))
            only = Wp.homo_warping_3D(cu(feats[:, v]), cu(src_p), cu(ref_p), cu(dv))
            assert torch.equal(warped, only)
            gw = torch.from_numpy(g["warped_%s_v%d" % (tag, v)])
            gm = torch.from_numpy(g["mask_%s_v%d" % (tag, v)])
            assert rel_l1(warped.cpu(), gw) < 2e-5
            assert (mask.cpu() != gm).float().mean() < 2e-3


def _check_similarity(net, sim, sim_o, feats, cams, hyp, tol, agree):
    """The channels-last kernels return sim_depth [B,H,W] (argmax taken inside the kernel): compare the chosen hypotheses with
    the oracle's argmax, and hold the similarity VOLUME of the NCHW kernels (same arithmetic) to the oracle's."""
    from mvsformer_b200 import config
    if sim.dim() == 3:
        want = torch.gather(hyp, 1, sim_o.argmax(dim=1, keepdim=True)).squeeze(1)
        assert (sim.cpu() == want).float().mean() > agree
        config.set_cv_layout("nchw")
        try:
            sim = net.build_cost_volume(cu(feats), cu(cams), cu(hyp))[1]
        finally:
            config.set_cv_layout("cl")
    assert rel_l1(sim.cpu(), sim_o) < tol


@pytest.mark.parametrize("s", [0, 1, 2, 3])
@pytest.mark.parametrize("smooth", [True, False])
def test_cost_volume_parts(s, smooth):
    """pass A (entropy, cosine similarity), vis net, pass B (aggregated volume) vs the oracle."""
    height, width, batch, views = 128, 192, 2, 4
    feats = S.make_features(batch, views, height, width, stages=(s,), smooth=smooth)["stage%d" % (s + 1)]
    cams = S.make_cameras(batch, views, height, width)["stage%d" % (s + 1)]
    hyp = S.narrow_hypotheses(s, height, width, batch)
    net = StageNet(dict(STAGE_ARGS), S.NDEPTHS[s], s).eval()
    sd = S.fill_state_dict(net.state_dict(), seed=s)
    net.load_state_dict(sd)
    net = net.to(DEV)
    vol_o, sim_o, parts = O.build_cost_volume(feats, cams, hyp, sd, want_parts=True)
    volume, sim, entropy, weight = net.build_cost_volume(cu(feats), cu(cams), cu(hyp))
    tol = 2e-5 if smooth else 3e-4
    ent_o = torch.cat(parts["entropy"], dim=1)
    wgt_o = torch.cat(parts["weight"], dim=1)
    assert rel_l1(entropy.cpu(), ent_o) < tol
    assert rel_l1(weight.cpu(), wgt_o) < tol
    _check_similarity(net, sim, sim_o, feats, cams, hyp, tol, 0.99 if smooth else 0.97)
    assert rel_l1(volume.cpu().permute(0, 4, 1, 2, 3), vol_o) < tol
    # the vis net alone, on the oracle's entropy (isolates the fused 2D CNN)
    b, n, h, w = ent_o.shape
    w_only = engine.vis_weight(cu(ent_o.reshape(b * n, h, w).contiguous()), net._vis_params_host()).view(b, n, h, w)
    assert rel_l1(w_only.cpu(), wgt_o) < 1e-5


@pytest.mark.parametrize("s", [0, 3])
def test_cost_volume_wild_geometry(s):
    """Source footprints far larger than the TMA box (strong rotation + roll, hypotheses sweeping the
    whole range, some samples behind the camera): every sample must still match the reference's
    per-tap zero-padding semantics (the kernels fall back to their predicated global path)."""
    height, width, batch, views = 128, 192, 1, 3
    feats = S.make_features(batch, views, height, width, stages=(s,), seed=99)["stage%d" % (s + 1)]
    cams = S.make_cameras(batch, views, height, width)["stage%d" % (s + 1)].clone()
    import math
    for v, (th, roll) in enumerate([(0.0, 0.0), (0.9, 0.5), (-0.6, -1.2)]):
        c, sn, cr, sr = math.cos(th), math.sin(th), math.cos(roll), math.sin(roll)
        ry = torch.tensor([[c, 0, sn], [0, 1, 0], [-sn, 0, c]], dtype=torch.float32)
        rz = torch.tensor([[cr, -sr, 0], [sr, cr, 0], [0, 0, 1]], dtype=torch.float32)
        rot = rz @ ry
        cams[0, v, 0, :3, :3] = rot
        cams[0, v, 0, :3, 3] = torch.tensor([0.0, 0.0, 680.0]) - rot @ torch.tensor([0.0, 0.0, 680.0]) + torch.tensor([40.0 * v, 0.0, -300.0 * v])
    nd = S.NDEPTHS[s]
    h, w = S.stage_hw(height, width, s)
    hyp = torch.linspace(150.0, 2500.0, nd).view(1, nd, 1, 1).repeat(1, 1, h, w).contiguous()
    net = StageNet(dict(STAGE_ARGS), nd, s).eval()
    sd = S.fill_state_dict(net.state_dict(), seed=s)
    net.load_state_dict(sd)
    net = net.to(DEV)
    vol_o, sim_o, parts = O.build_cost_volume(feats, cams, hyp, sd, want_parts=True)
    volume, sim, entropy, weight = net.build_cost_volume(cu(feats), cu(cams), cu(hyp))
    assert rel_l1(entropy.cpu(), torch.cat(parts["entropy"], dim=1)) < 1e-4
    _check_similarity(net, sim, sim_o, feats, cams, hyp, 1e-4, 0.97)
    assert rel_l1(volume.cpu().permute(0, 4, 1, 2, 3), vol_o) < 1e-4


@pytest.mark.parametrize("s", [0, 1, 2, 3])
def test_cost_volume_is_deterministic_under_allocator_churn(s):
    """The cost-volume build must give bit-identical results call after call, also when every output
    and scratch buffer lands on recycled memory that holds NaNs (no read of uninitialised memory, no
    race between the TMA-staged tiles and their consumers)."""
    height, width, batch, views = 128, 192, 2, 4
    feats = cu(S.make_features(batch, views, height, width, stages=(s,))["stage%d" % (s + 1)])
    cams = cu(S.make_cameras(batch, views, height, width)["stage%d" % (s + 1)])
    hyp = cu(S.narrow_hypotheses(s, height, width, batch))
    net = StageNet(dict(STAGE_ARGS), S.NDEPTHS[s], s).eval()
    net.load_state_dict(S.fill_state_dict(net.state_dict(), seed=s))
    net = net.to(DEV)
    first = None
    gen = torch.Generator().manual_seed(5)
    for it in range(12):
        # poison the caching allocator's free blocks with NaNs of assorted sizes
        junk = [torch.full((int(n),), float("nan"), device=DEV) for n in torch.randint(1 << 8, 1 << 20, (6,), generator=gen)]
        del junk
        out = [t.clone() for t in net.build_cost_volume(feats, cams, hyp)]
        assert all(torch.isfinite(t).all() for t in out)
        if first is None:
            first = out
        else:
            for a, b in zip(out, first):
                assert torch.equal(a, b), "iteration %d differs from the first call" % it


CONV_CASES = [
    # cin, cout, kd, stride, D, H, W
    (8, 16, 3, (2, 2, 2), 8, 16, 24),
    (8, 16, 3, (1, 2, 2), 4, 16, 24),
    (16, 16, 3, (1, 1, 1), 4, 10, 14),
    (16, 32, 3, (2, 2, 2), 8, 8, 12),
    (32, 32, 3, (1, 1, 1), 2, 6, 10),
    (32, 64, 3, (1, 2, 2), 3, 6, 10),
    (64, 64, 3, (1, 1, 1), 2, 5, 7),
    (8, 16, 1, (1, 2, 2), 3, 16, 24),
]


@pytest.mark.parametrize("cin,cout,kd,stride,D,H,W", CONV_CASES)
def test_conv3d_block(cin, cout, kd, stride, D, H, W):
    g = S._gen(cin * 100 + cout + kd)
    blk = M.Conv3d(cin, cout, kernel_size=(kd, 3, 3), stride=stride, padding=(kd // 2, 1, 1)).eval()
    blk.load_state_dict(S.fill_state_dict(blk.state_dict(), seed=3))
    x = torch.randn(2, cin, D, H, W, generator=g)
    sd = blk.state_dict()
    want = F.conv3d(x, sd["conv.weight"], stride=stride, padding=(kd // 2, 1, 1))
    want = torch.relu(O._bn_eval(want, sd, "bn"))
    got = blk.to(DEV)(cu(x)).cpu()
    assert got.shape == want.shape
    assert rel_l1(got, want) < 1e-5          # default conv precision is 3xTF32 on the tensor cores
    # skip connection is added AFTER the activation
    skip = torch.randn(want.shape, generator=g)
    got2 = engine.cl_to_ncdhw(blk.forward_cl(engine.ncdhw_to_cl(cu(x)), skip=engine.ncdhw_to_cl(cu(skip)))).cpu()
    assert rel_l1(got2, want + skip) < 1e-5


DECONV_CASES = [
    (64, 32, 3, 2, 2, 3, 5), (32, 16, 3, 2, 3, 6, 10), (16, 8, 3, 2, 4, 8, 12),
    (64, 32, 3, 1, 2, 3, 5), (32, 16, 3, 1, 4, 6, 10), (16, 8, 3, 1, 4, 8, 12), (16, 8, 1, 1, 3, 8, 12),
]


@pytest.mark.parametrize("cin,cout,kd,sd,D,H,W", DECONV_CASES)
def test_deconv3d_block(cin, cout, kd, sd, D, H, W):
    g = S._gen(cin * 7 + cout + kd + sd)
    blk = M._SeqDeconv(cin, cout, (kd, 3, 3), (sd, 2, 2), (kd // 2, 1, 1), (sd - 1, 1, 1)).eval()
    blk.load_state_dict(S.fill_state_dict(blk.state_dict(), seed=4))
    sdict = blk.state_dict()
    x = torch.randn(2, cin, D, H, W, generator=g)
    want = F.conv_transpose3d(x, sdict["0.weight"], stride=(sd, 2, 2), padding=(kd // 2, 1, 1), output_padding=(sd - 1, 1, 1))
    want = torch.relu(O._bn_eval(want, sdict, "1"))
    got = blk.to(DEV)(cu(x)).cpu()
    assert got.shape == want.shape
    assert rel_l1(got, want) < 1e-5


@pytest.mark.parametrize("kind", ["CostRegNet", "CostRegNet3D", "CostRegNet2D"])
def test_cost_reg_nets(kind):
    g = S._gen(99)
    net = getattr(M, kind)(8, 8).eval()
    sd = S.fill_state_dict(net.state_dict(), seed=6)
    net.load_state_dict(sd)
    x = torch.randn(2, 8, 8, 16, 24, generator=g)
    fn = {"CostRegNet": O.cost_reg_net, "CostRegNet3D": O.cost_reg_net_3d, "CostRegNet2D": O.cost_reg_net_2d}[kind]
    want = fn(x, {"cost_reg." + k: v for k, v in sd.items()})
    got = net.to(DEV)(cu(x)).cpu()
    assert got.shape == want.shape == (2, 1, 8, 16, 24)
    assert rel_l1(got, want) < 2e-5


def test_cost_reg_rejects_bad_shapes():
    net = M.CostRegNet(8, 8).eval().to(DEV)
    with pytest.raises(RuntimeError):
        net(torch.zeros(1, 8, 8, 12, 24, device=DEV))          # H not divisible by 8 (reference: size mismatch)
    with pytest.raises(RuntimeError):
        M.depth_regression(torch.zeros(1, 4, 8, 8), torch.zeros(1, 4))   # CPU tensors: no fallback


def test_schedules_vs_golden():
    g = load_golden("schedules.npz")
    dv = cu(S.make_depth_range(2))
    assert max_abs(M.init_inverse_range(dv, 32, DEV, torch.float32, 8, 12).cpu(), g["init_inverse_range"]) < 2e-4
    assert max_abs(M.init_range(dv, 32, DEV, torch.float32, 8, 12).cpu(), g["init_range"]) < 2e-4
    hyp = cu(torch.from_numpy(g["init_inverse_range"]))
    depth = cu(torch.from_numpy(g["sched_depth_in"]))
    assert rel_l1(M.schedule_inverse_range(depth, hyp, 16, 2.67, 16, 24).cpu(), g["schedule_inverse_range"]) < 1e-6
    assert rel_l1(M.schedule_range(depth, 16, 2.67 * (dv[:, 1] - dv[:, 0]), 16, 24).cpu(), g["schedule_range"]) < 1e-6
    p = cu(torch.from_numpy(g["prob_in"]))
    assert rel_l1(M.depth_regression(p, hyp).cpu(), g["depth_regression_map"]) < 1e-6
    assert rel_l1(M.depth_regression(p, dv[:, :32].contiguous()).cpu(), g["depth_regression_vec"]) < 1e-6
    for n in (2, 3, 4):
        assert rel_l1(M.conf_regression(p, n).cpu(), g["conf_regression_n%d" % n]) < 1e-6


def test_regression_head_modes():
    g = S._gen(21)
    pre = 2.0 * torch.randn(2, 16, 12, 20, generator=g)
    dv = S.narrow_hypotheses(1, 48, 80, 2)
    for training in (False, True):
        prob_o, depth_o, conf_o = O.regression_head(pre, dv, 5.0, training)
        prob, depth, conf = engine.regression_head(cu(pre), cu(dv), 5.0, training)
        assert rel_l1(prob.cpu(), prob_o) < 1e-6
        assert rel_l1(conf.cpu(), conf_o) < 1e-6
        if training:
            assert torch.equal(depth.cpu(), depth_o)
        else:
            assert rel_l1(depth.cpu(), depth_o) < 1e-6


@pytest.mark.parametrize("s", [0, 1, 2, 3])
def test_stagenet_vs_golden(s):
    g = load_golden("stage%d.npz" % (s + 1))
    batch = int(g["batch"])
    height, width = 128, 192
    feats = S.make_features(batch, 3, height, width, stages=(s,))["stage%d" % (s + 1)]
    assert checksum(feats) == pytest.approx(float(g["feat_checksum"]), rel=1e-12)
    cams = S.make_cameras(batch, 3, height, width)["stage%d" % (s + 1)]
    hyp = S.narrow_hypotheses(s, height, width, batch)
    net = StageNet(dict(STAGE_ARGS), S.NDEPTHS[s], s).eval()
    net.load_state_dict(S.fill_state_dict(net.state_dict(), seed=s), strict=True)
    out = net.to(DEV)(cu(feats), cu(cams), cu(hyp), tmp=list(S.EVAL_TMP))
    assert set(out) == {"depth", "prob_volume", "photometric_confidence", "depth_values", "prob_volume_pre", "sim_depth"}
    assert rel_l1(out["prob_volume_pre"].cpu(), g["prob_volume_pre"]) < 2e-5
    assert rel_l1(out["prob_volume"].cpu(), g["prob_volume"]) < 2e-5
    assert rel_l1(out["depth"].cpu(), g["depth"]) < 1e-5
    assert rel_l1(out["photometric_confidence"].cpu(), g["photometric_confidence"]) < 2e-5
    # argmax over a similarity volume with exact ties in fp32 (the stage-1 case has top-1 == top-2 pixels): a different,
    # equally valid summation order over the channels flips a handful of them (cosine sums agree to 1.2e-7 absolute,
    # measured on B200: 3 of 768 pixels differ for the round-1 kernels, 4 for the channels-last ones)
    assert (out["sim_depth"].cpu() == torch.from_numpy(g["sim_depth"])).float().mean() > 0.99


def test_cascade_vs_golden():
    g = load_golden("cascade.npz")
    height, width, batch, views = int(g["height"]), int(g["width"]), int(g["batch"]), int(g["views"])
    feats = S.make_features(batch, views, height, width, seed=int(g["feat_seed"]))
    cams = S.make_cameras(batch, views, height, width)
    dv = S.make_depth_range(batch)
    args = dict(STAGE_ARGS, ndepths=list(S.NDEPTHS), depth_interals_ratio=list(S.DEPTH_INTERVAL_RATIO), inverse_depth=True)
    net = CascadeMVS(args).eval()
    full = {}
    for s in range(4):
        sd = S.fill_state_dict(net.fusions[s].state_dict(), seed=int(g["weight_seed0"]) + s)
        full.update({"fusions.%d.%s" % (s, k): v for k, v in sd.items()})
    net.load_state_dict(full, strict=True)
    net = net.to(DEV)
    out = net({k: cu(v) for k, v in feats.items()}, {k: cu(v) for k, v in cams.items()}, cu(dv), tmp=list(S.EVAL_TMP))
    for s in range(4):
        st = out["stage%d" % (s + 1)]
        assert rel_l1(st["depth_values"].cpu(), g["stage%d_depth_values" % (s + 1)]) < 1e-5
        assert rel_l1(st["depth"].cpu(), g["stage%d_depth" % (s + 1)]) < 1e-5
        assert rel_l1(st["prob_volume_pre"].cpu(), g["stage%d_prob_volume_pre" % (s + 1)]) < 1e-4
    # north-star bar is 1e-3 relative L1 on the depth map
    assert rel_l1(out["refined_depth"].cpu(), g["refined_depth"]) < 1e-5
    assert rel_l1(out["photometric_confidence"].cpu(), g["photometric_confidence"]) < 2e-5


def test_full_size_properties():
    """BASELINE cfg 2 size (1152x1536, 5 views): size-independent properties of the cascade."""
    height, width, batch, views = 1152, 1536, 1, 5
    feats = {k: cu(v) for k, v in S.make_features(batch, views, height, width).items()}
    cams = {k: cu(v) for k, v in S.make_cameras(batch, views, height, width).items()}
    dv = cu(S.make_depth_range(batch))
    args = dict(STAGE_ARGS, ndepths=list(S.NDEPTHS), depth_interals_ratio=list(S.DEPTH_INTERVAL_RATIO), inverse_depth=True)
    net = CascadeMVS(args).eval()
    full = {}
    for s in range(4):
        sd = S.fill_state_dict(net.fusions[s].state_dict(), seed=s)
        full.update({"fusions.%d.%s" % (s, k): v for k, v in sd.items()})
    net.load_state_dict(full)
    net = net.to(DEV)
    out = net(feats, cams, dv, tmp=list(S.EVAL_TMP))
    torch.cuda.synchronize()
    for s in range(4):
        st = out["stage%d" % (s + 1)]
        hyp = st["depth_values"]
        assert torch.isfinite(st["prob_volume_pre"]).all()
        # regression is a convex combination of the hypotheses
        assert (st["depth"] <= hyp.max(dim=1)[0] * (1 + 1e-6)).all() and (st["depth"] >= hyp.min(dim=1)[0] * (1 - 1e-6)).all()
        # softmax sums to one; confidence is its maximum
        assert (st["prob_volume"].sum(dim=1) - 1).abs().max() < 1e-5
        # hypotheses are monotone (inverse-depth schedule: far -> near)
        assert (hyp[:, 1:] < hyp[:, :-1]).all()
    assert out["refined_depth"].shape == (1, height, width)
    conf = out["photometric_confidence"]
    assert (conf > 0).all() and (conf <= 1 + 1e-6).all()
    # permuting the source views must not change the aggregated volume (sum over views)
    s = 3
    net4 = net.fusions[s]
    f4, c4, h4 = feats["stage4"], cams["stage4"], out["stage4"]["depth_values"]
    perm = [0, 3, 1, 4, 2]
    v1 = net4.build_cost_volume(f4, c4, h4)[0]
    v2 = net4.build_cost_volume(f4[:, perm].contiguous(), c4[:, perm].contiguous(), h4)[0]
    assert rel_l1(v2, v1) < 1e-6


@pytest.mark.parametrize("batch,views,height,width", [(1, 2, 64, 96), (2, 11, 64, 128), (1, 4, 64, 104)])
def test_stagenet_odd_shapes_vs_oracle(batch, views, height, width):
    """Single source view, 10 source views (T&T), batch 2, and a width whose stage-4 row is not a
    multiple of 32 / TMA-friendly (generic kernels): every stage against the oracle."""
    for s in (0, 3):
        h, w = S.stage_hw(height, width, s)
        if s == 0 and (h % 8 or w % 8):
            continue
        feats = S.make_features(batch, views, height, width, stages=(s,), seed=31 + s)["stage%d" % (s + 1)]
        cams = S.make_cameras(batch, views, height, width)["stage%d" % (s + 1)]
        hyp = S.narrow_hypotheses(s, height, width, batch)
        net = StageNet(dict(STAGE_ARGS), S.NDEPTHS[s], s).eval()
        sd = S.fill_state_dict(net.state_dict(), seed=s)
        net.load_state_dict(sd)
        want = O.stage_forward(feats, cams, hyp, sd, S.NDEPTHS[s], S.EVAL_TMP[s])
        out = net.to(DEV)(cu(feats), cu(cams), cu(hyp), tmp=list(S.EVAL_TMP))
        assert rel_l1(out["prob_volume_pre"].cpu(), want["prob_volume_pre"]) < 5e-5
        assert rel_l1(out["depth"].cpu(), want["depth"]) < 1e-5
        assert rel_l1(out["photometric_confidence"].cpu(), want["photometric_confidence"]) < 5e-5


def test_half_precision_features_are_upcast():
    """AMP callers hand fp16 features (mvsformer_model.py:68,78 upcasts to fp32 inside autocast(False))."""
    s, height, width = 3, 64, 96
    feats = S.make_features(1, 3, height, width, stages=(s,))["stage4"].half()
    cams = S.make_cameras(1, 3, height, width)["stage4"]
    hyp = S.narrow_hypotheses(s, height, width, 1)
    net = StageNet(dict(STAGE_ARGS), S.NDEPTHS[s], s).eval()
    sd = S.fill_state_dict(net.state_dict(), seed=s)
    net.load_state_dict(sd)
    want = O.stage_forward(feats.float(), cams, hyp, sd, S.NDEPTHS[s], S.EVAL_TMP[s])
    out = net.to(DEV)(cu(feats), cu(cams), cu(hyp), tmp=list(S.EVAL_TMP))
    assert rel_l1(out["depth"].cpu(), want["depth"]) < 1e-5


def test_non_contiguous_feature_views():
    """features assembled by torch.stack of per-view tensors with a non-dense batch stride fall back to
    the generic (non-TMA) kernels and give the same volume."""
    s, height, width = 2, 64, 96
    feats = S.make_features(2, 3, height, width, stages=(s,))["stage3"]
    padded = torch.zeros(2, 4, *feats.shape[2:])
    padded[:, :3] = feats
    view = padded.to(DEV)[:, :3]                      # batch stride = 4*C*h*w, view stride = C*h*w
    cams = S.make_cameras(2, 3, height, width)["stage3"]
    hyp = S.narrow_hypotheses(s, height, width, 2)
    net = StageNet(dict(STAGE_ARGS), S.NDEPTHS[s], s).eval().to(DEV)
    a = net.build_cost_volume(view, cu(cams), cu(hyp))[0]
    b = net.build_cost_volume(cu(feats), cu(cams), cu(hyp))[0]
    assert torch.equal(a, b)


def _cascade(seed0=0):
    args = dict(STAGE_ARGS, ndepths=list(S.NDEPTHS), depth_interals_ratio=list(S.DEPTH_INTERVAL_RATIO), inverse_depth=True)
    net = CascadeMVS(args).eval()
    full, sds = {}, []
    for s in range(4):
        sd = S.fill_state_dict(net.fusions[s].state_dict(), seed=seed0 + s)
        sds.append(sd)
        full.update({"fusions.%d.%s" % (s, k): v for k, v in sd.items()})
    net.load_state_dict(full, strict=True)
    return net.to(DEV), sds


def test_baseline_cfg1_plumbing_case():
    """BASELINE.json configs[0]: 2 source views, 32 hypotheses, 320x256 image, single cascade stage
    (stage-1 features [1,3,64,32,40]) — the reference's CPU-runnable case, against the oracle."""
    height, width = 256, 320
    feats = S.make_features(1, 3, height, width, stages=(0,), seed=2024)["stage1"]
    cams = S.make_cameras(1, 3, height, width)["stage1"]
    dv = S.make_depth_range(1)
    hyp = O.init_inverse_range(dv, 32, 32, 40)
    net = StageNet(dict(STAGE_ARGS), 32, 0).eval()
    sd = S.fill_state_dict(net.state_dict(), seed=0)
    net.load_state_dict(sd)
    want = O.stage_forward(feats, cams, hyp, sd, 32, 5.0)
    got_hyp = M.init_inverse_range(cu(dv), 32, DEV, torch.float32, 32, 40)
    assert max_abs(got_hyp.cpu(), hyp) < 2e-4
    out = net.to(DEV)(cu(feats), cu(cams), cu(hyp), tmp=5.0)       # same hypothesis values on both sides (sim_depth is an exact gather)
    assert rel_l1(out["depth"].cpu(), want["depth"]) < 1e-5
    assert rel_l1(out["prob_volume"].cpu(), want["prob_volume"]) < 5e-5
    assert (out["sim_depth"].cpu() == want["sim_depth"]).float().mean() > 0.99


@pytest.mark.parametrize("name,batch,views,height,width", [("cfg3 BlendedMVS", 4, 7, 576, 768), ("cfg4 T&T @1088", 1, 11, 1088, 1920)])
def test_baseline_cfg3_cfg4_properties(name, batch, views, height, width):
    """BASELINE.json configs[2] (B=4, 7 views, 576x768) and configs[3] (11 views, 1088x1920: 1056 is not
    runnable by the reference): size-independent properties of the full cascade."""
    net, _ = _cascade()
    feats = {k: cu(v) for k, v in S.make_features(batch, views, height, width, seed=9).items()}
    cams = {k: cu(v) for k, v in S.make_cameras(batch, views, height, width).items()}
    dv = cu(S.make_depth_range(batch))
    out = net(feats, cams, dv, tmp=list(S.EVAL_TMP))
    torch.cuda.synchronize()
    assert out["refined_depth"].shape == (batch, height, width)
    for s in range(4):
        st = out["stage%d" % (s + 1)]
        hyp = st["depth_values"]
        assert torch.isfinite(st["prob_volume_pre"]).all()
        assert (st["depth"] <= hyp.max(dim=1)[0] * (1 + 1e-6)).all() and (st["depth"] >= hyp.min(dim=1)[0] * (1 - 1e-6)).all()
        assert (st["prob_volume"].sum(dim=1) - 1).abs().max() < 1e-5
    conf = out["photometric_confidence"]
    assert (conf > 0).all() and (conf <= 1 + 1e-6).all()
    # batch items are independent: item 0 alone reproduces item 0 of the batch
    if batch > 1:
        one = net({k: v[:1].contiguous() for k, v in feats.items()}, {k: v[:1].contiguous() for k, v in cams.items()},
                  dv[:1].contiguous(), tmp=list(S.EVAL_TMP))
        assert rel_l1(one["refined_depth"], out["refined_depth"][:1]) < 1e-6


@pytest.mark.parametrize("shape", [(2, 5, 13, 37), (1, 17, 9, 33), (1, 1, 3, 3), (1, 8, 8, 32), (1, 32, 72, 96), (3, 9, 16, 70)])
@pytest.mark.parametrize("ksize", [1, 3])
def test_prob_conv_entry_point(shape, ksize):
    """mvs_prob_conv_cl vs fp64 F.conv3d: the 3x3x3 kernel marches a 32x8 pixel tile over 8-slice chunks (partial tiles,
    volumes thinner than a chunk, chunk seams at z = 8, 16, ...); the 1x1x1 one adds the bias."""
    b, d, h, w = shape
    g = S._gen(b * 1000 + d * 100 + h + w + ksize)
    x = torch.randn(b, d, h, w, 8, generator=g)
    wt = torch.randn(ksize ** 3, 8, generator=g) * 0.3
    bias = torch.randn(1, generator=g) if ksize == 1 else None
    w5 = wt.reshape(ksize, ksize, ksize, 8).permute(3, 0, 1, 2).unsqueeze(0).double()
    want = F.conv3d(x.permute(0, 4, 1, 2, 3).double(), w5, bias.double() if bias is not None else None, padding=ksize // 2)[:, 0]
    got = engine.prob_conv_cl(cu(x), np.ascontiguousarray(wt.numpy()), bias.numpy() if bias is not None else None, ksize).cpu()
    assert got.shape == want.shape
    assert rel_l1(got, want) < 2e-6 and max_abs(got, want.float()) < 2e-5
