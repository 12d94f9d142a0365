"""GPU tests of the two implementations behind each ``mvsformer_b200.config`` switch (channels-last vs generic cost-volume
kernels, fused vs CUDA-core visibility net) and of the paths next to the hot path (depth-map fusion, the heads / fusion types
no shipped config uses, diff-warp gradients, tensor-core training convolutions).  The round-1 "opt-in variants" that lost
the A/B on the B200 (profiles/r02_ab_variants.json, profiles/r02_gated_tests_first_run.log) or were superseded by the
round-2 kernels have been deleted together with their tests.
"""
import os

import pytest
import torch

from mvsformer_b200 import config, synthetic as S
from mvsformer_b200.mvsformer_model import StageNet
from tests.helpers import STAGE_ARGS

pytestmark = [pytest.mark.gpu]
DEV = "cuda"


def test_fusion_gpu_vs_reference_golden_and_oracle():
    """Depth-map fusion kernels (csrc/fusion.cu) on the device vs the reference's recorded outputs and the fp64 oracle
    (the same kernel source passes tests/test_fusion.py on the CPU emulation)."""
    from mvsformer_b200 import fusion as Fu
    from oracle import fusion_oracle as FO
    from tests.helpers import load_golden, rel_l1

    g = load_golden("fusion.npz")
    c = S.make_fusion_case(int(g["views"]), int(g["height"]), int(g["width"]), seed=int(g["seed"]))
    d = {k: v.to(DEV) for k, v in c.items()}
    out = Fu.filter_view(d["ref_depth"], d["src_depths"], d["ref_cam"], d["src_cams"], 1.0, 0.01, 2,
                         ref_conf=d["ref_conf"], prob_thresh=[0.1, 0.2, 0.3])
    torch.cuda.synchronize()
    agree = lambda a, b: float((a.cpu().bool() == torch.as_tensor(b).bool()).float().mean())
    assert agree(out["in_range"], g["in_range"]) > 0.995
    assert agree(out["vis_masks"], g["masks"]) > 0.99 and agree(out["vis_mask"], g["mask"]) > 0.99
    both = (torch.from_numpy(g["in_range"]) > 0.5) & (out["in_range"].cpu() > 0.5)
    sel = both.expand(-1, -1, 3, -1, -1)
    assert rel_l1(out["reproj_xyd"].cpu()[sel], torch.from_numpy(g["reproj_xyd"])[sel]) < 1e-4
    c64 = {k: v.double() for k, v in c.items()}
    xyd64, _ = FO.get_reproj(c64["ref_depth"], c64["src_depths"], c64["ref_cam"], c64["src_cams"])
    assert rel_l1(out["reproj_xyd"].cpu()[sel], xyd64[sel]) < 1e-4
    same = (out["vis_masks"].cpu() == torch.from_numpy(g["masks"])).all(dim=1)
    assert rel_l1(out["depth_ave"].cpu()[same], torch.from_numpy(g["ave"])[same]) < 1e-5
    # full DTU size, 10 source views (test.py:405): properties only
    big = S.make_fusion_case(11, 1152, 1536, seed=2, outlier_frac=0.02)
    bd = {k: v.to(DEV) for k, v in big.items()}
    res = Fu.filter_view(bd["ref_depth"], bd["src_depths"], bd["ref_cam"], bd["src_cams"], 1.0, 0.01, 3)
    torch.cuda.synchronize()
    frac = float(res["vis_mask"].float().mean())
    assert 0.5 < frac <= 1.0                                   # a consistent plane: most pixels survive
    kept = res["vis_mask"]
    assert float((res["depth_ave"][kept] - bd["ref_depth"][kept]).abs().max()) < 0.02 * float(bd["ref_depth"].max())


@pytest.mark.parametrize("kind", ["mixup_ce", "re"])
@pytest.mark.parametrize("s", [1, 3])
def test_other_depth_type_heads_gpu(kind, s):
    """StageNet with depth_type 'mixup_ce' / 're' on the device vs the unmodified reference (tests/golden/heads.npz)."""
    from tests.helpers import load_golden, rel_l1
    g = load_golden("heads.npz")
    height, width = int(g["height"]), int(g["width"])
    net = StageNet(dict(STAGE_ARGS, depth_type=kind), S.NDEPTHS[s], s).eval()
    net.load_state_dict(S.fill_state_dict(net.state_dict(), seed=70 + s))
    net = net.to(DEV)
    feats = S.make_features(1, 3, height, width, seed=90 + s, stages=(s,))["stage%d" % (s + 1)]
    cams = S.make_cameras(1, 3, height, width)["stage%d" % (s + 1)]
    hyp = S.narrow_hypotheses(s, height, width, 1)
    with torch.no_grad():
        out = net(feats.to(DEV), cams.to(DEV), hyp.to(DEV), tmp=list(S.EVAL_TMP))
    assert rel_l1(out["prob_volume"].cpu(), g["%s_s%d_prob_volume" % (kind, s + 1)]) < 1e-4
    assert rel_l1(out["depth"].cpu(), g["%s_s%d_depth" % (kind, s + 1)]) < 1e-4
    assert rel_l1(out["photometric_confidence"].cpu(), g["%s_s%d_photometric_confidence" % (kind, s + 1)]) < 1e-4


def test_dynamic_fusion_gpu_vs_reference_golden():
    from mvsformer_b200 import fusion as Fu
    from tests.helpers import load_golden, rel_l1

    g = load_golden("fusion.npz")
    c = S.make_fusion_case(int(g["views"]), int(g["height"]), int(g["width"]), seed=int(g["seed"]))
    d = {k: v.to(DEV) for k, v in c.items()}
    out = Fu.dynamic_filter_view(d["ref_depth"], d["src_depths"], d["ref_cam"], d["src_cams"], 4, 1300)
    torch.cuda.synchronize()
    gx = torch.from_numpy(g["dyn_xyd"])
    xyd = out["reproj_xyd"].cpu()
    finite = torch.isfinite(gx) & torch.isfinite(xyd)
    assert rel_l1(xyd[finite], gx[finite]) < 1e-4
    agree = lambda a, b: float((a.cpu().bool() == torch.as_tensor(b).bool()).float().mean())
    assert agree(out["vis_mask"], g["dyn_mask"]) > 0.99 and agree(out["geo_mask"], g["dyn_geo_mask"]) > 0.99
    levels, mask = Fu.vis_filter_dynamic(d["ref_depth"], gx.to(DEV), 4, 1300)
    assert torch.equal(levels.cpu(), torch.from_numpy(g["dyn_level_counts"])) and agree(mask, g["dyn_mask"]) == 1.0


@pytest.mark.parametrize("mode,tol", [("tf32x3", 5e-3), ("tf32", 5e-2)])
def test_training_convs_on_tensor_cores(mode, tol):
    """MVS_TRAIN_CONV: forward / data-gradient convolutions of the training path through the tcgen05 kernels;
    gradients vs the fp64 oracle (tf32x3 at the fp32 path's tolerance, tf32 loosely)."""
    import torch.nn.functional as F
    from tests.helpers import rel_l1
    from tests.test_gpu_train import _oracle_gradients
    from tests.test_oracle_golden import _train_grad_case

    s = 1
    g, feats, cams, hyp, sd, target = _train_grad_case(s)
    net = StageNet(dict(STAGE_ARGS), S.NDEPTHS[s], s).train()
    net.load_state_dict(sd)
    net = net.to(DEV)
    config.set_train_conv(mode)
    try:
        f1 = feats.to(DEV).requires_grad_(True)
        out = net(f1, cams.to(DEV), hyp.to(DEV))
        F.cross_entropy(out["prob_volume_pre"], target.to(DEV)).backward()
        torch.cuda.synchronize()
    finally:
        config.set_train_conv("fp32")
    truth = _oracle_gradients(s, torch.float64)
    assert rel_l1(out["prob_volume_pre"].cpu(), g["prob_volume_pre"]) < tol
    assert rel_l1(f1.grad.cpu(), truth["features"]) < 2 * tol
    for name, p in net.named_parameters():
        if name == "cost_reg.prob.bias" or name.startswith("vis."):
            continue
        # parameter gradients are heavily cancelling sums: same bar as tests/test_gpu_train.py (the reference's own
        # fp32 arithmetic is up to 7e-3 away from the fp64 value for cost_reg.conv1); measured 1.07e-2 for tf32x3
        assert rel_l1(p.grad.cpu(), truth[name]) < 4 * tol, name


def test_diff_homo_warping_gradients_gpu():
    """diff_homo_warping_3D_with_mask on the device: gradients w.r.t. features, depth and both cameras vs torch autograd
    through the reference's formula (the kernel source passes the same check on the CPU emulation)."""
    import torch.nn.functional as F
    from mvsformer_b200 import warping as Wp
    from oracle import mvs_oracle as O
    from tests.helpers import rel_l1
    from tests.test_train_emulated import _case

    feats, cams, hyp = _case(batch=2, views=2, chans=8, depth=4, height=16, width=24, seed=23)
    cams = cams.clone()
    cams[:, 1, 0, 0, 3] += 120.0
    src_p, ref_p = O.compose_projection(cams[:, 1]), O.compose_projection(cams[:, 0])
    cpu = [t.clone().requires_grad_(True) for t in (feats[:, 1], src_p, ref_p, hyp)]
    b, c, h, w = cpu[0].shape
    nd = hyp.shape[1]
    proj = torch.matmul(cpu[1], torch.inverse(cpu[2]))
    y, x = torch.meshgrid(torch.arange(h, dtype=torch.float32), torch.arange(w, dtype=torch.float32), indexing="ij")
    xyz = torch.stack((x.reshape(-1), y.reshape(-1), torch.ones(h * w))).unsqueeze(0).repeat(b, 1, 1)
    pxyz = torch.matmul(proj[:, :3, :3], xyz).unsqueeze(2).repeat(1, 1, nd, 1) * cpu[3].reshape(b, 1, nd, -1) + proj[:, :3, 3:4].view(b, 3, 1, 1)
    pxy = pxyz[:, :2] / (pxyz[:, 2:3] + 1e-6)
    grid = torch.stack((pxy[:, 0] / ((w - 1) / 2) - 1, pxy[:, 1] / ((h - 1) / 2) - 1), dim=3)
    want = F.grid_sample(cpu[0], grid.view(b, nd * h, w, 2), mode="bilinear", padding_mode="zeros", align_corners=True).view(b, c, nd, h, w)
    gout = torch.randn(want.shape, generator=S._gen(3))
    want.backward(gout)
    dev = [t.detach().to(DEV).requires_grad_(True) for t in (feats[:, 1].contiguous(), src_p, ref_p, hyp)]
    got, mask = Wp.diff_homo_warping_3D_with_mask(*dev)
    got.backward(gout.to(DEV))
    assert rel_l1(got.cpu(), want) < 1e-5
    for a, bb, tol in zip(dev, cpu, (1e-4, 2e-3, 2e-3, 2e-3)):
        assert rel_l1(a.grad.cpu(), bb.grad) < tol


@pytest.mark.parametrize("kind", ["epipole", "epipoleV2"])
def test_epipole_fusion_gpu_vs_reference_golden(kind):
    """fusion_type 'epipole' / 'epipoleV2' on the device: train-mode forward + gradients and eval outputs vs the unmodified
    reference (tests/golden/epipole.npz)."""
    import torch.nn.functional as F
    from tests.helpers import load_golden, rel_l1
    g = load_golden("epipole.npz")
    s, height, width = 2, int(g["height"]), int(g["width"])
    args = dict(STAGE_ARGS, fusion_type=kind, attn_temp=2.0)
    feats = S.make_features(1, 3, height, width, seed=60, stages=(s,))["stage%d" % (s + 1)]
    cams = S.make_cameras(1, 3, height, width)["stage%d" % (s + 1)].clone()
    cams[:, 2, 0, 0, 3] += 90.0
    hyp = S.narrow_hypotheses(s, height, width, 1)
    target = torch.randint(0, S.NDEPTHS[s], (1, feats.shape[-2], feats.shape[-1]), generator=S._gen(700 + s))
    net = StageNet(args, S.NDEPTHS[s], s).train()
    net.load_state_dict(S.fill_state_dict(net.state_dict(), seed=80))
    if kind == "epipoleV2":
        with torch.no_grad():
            net.attn_temp.fill_(1.7)
    net = net.to(DEV)
    f = feats.to(DEV).requires_grad_(True)
    out = net(f, cams.to(DEV), hyp.to(DEV), tmp=list(S.EVAL_TMP))
    assert rel_l1(out["prob_volume_pre"].cpu(), g[kind + "_train_pre"]) < 2e-4
    F.cross_entropy(out["prob_volume_pre"], target.to(DEV)).backward()
    assert rel_l1(f.grad.cpu(), g[kind + "_train_gfeat"]) < 2e-2
    if kind == "epipoleV2":
        assert float(net.attn_temp.grad) == pytest.approx(float(g[kind + "_train_gtemp"]), rel=5e-2)
    ev = StageNet(args, S.NDEPTHS[s], s).eval()
    ev.load_state_dict(S.fill_state_dict(ev.state_dict(), seed=80))
    ev = ev.to(DEV)
    with torch.no_grad():
        res = ev(feats.to(DEV), cams.to(DEV), hyp.to(DEV), tmp=list(S.EVAL_TMP))
    assert rel_l1(res["prob_volume_pre"].cpu(), g[kind + "_eval_prob_volume_pre"]) < 2e-4
    assert rel_l1(res["depth"].cpu(), g[kind + "_eval_depth"]) < 1e-4
    assert (res["sim_depth"].cpu() == torch.from_numpy(g[kind + "_eval_sim_depth"])).float().mean() > 0.99


# ---- channels-last cost-volume kernels (csrc/cost_volume_cl.cu) ------------------------------------------------
def test_features_to_cl_is_an_exact_permutation():
    from mvsformer_b200 import engine
    ts = [torch.randn(2, 3, c, h, w, device=DEV) for c, h, w in ((64, 16, 24), (32, 32, 48), (16, 37, 50), (8, 5, 7))]
    outs = engine.features_to_cl(ts)
    for t, o in zip(ts, outs):
        assert o.shape == (2, 3) + (t.shape[3], t.shape[4], t.shape[2])
        assert torch.equal(o, t.permute(0, 1, 3, 4, 2).contiguous())


@pytest.mark.parametrize("s", [0, 1, 2, 3])
@pytest.mark.parametrize("wild", [False, True])
@pytest.mark.parametrize("shape", [(128, 192, 2, 4), (64, 104, 1, 3)])
def test_channels_last_cost_volume_matches_generic_kernels(s, wild, shape):
    """Same arithmetic, different summation order over the channels of a sample: volume / entropy / similarity of the
    channels-last kernels vs the generic NCHW kernels of csrc/cost_volume.cu (each separately pinned to the oracle in test_gpu_parity.py), incl.
    samples outside the staged box / the image (predicated global path) and partial tiles (width 104/8 = 13 px)."""
    from tests.helpers import rel_l1
    height, width, batch, views = shape
    feats = S.make_features(batch, views, height, width, stages=(s,), smooth=not wild)["stage%d" % (s + 1)]
    cams = S.make_cameras(batch, views, height, width)["stage%d" % (s + 1)].clone()
    hyp = S.narrow_hypotheses(s, height, width, batch)
    if wild:
        cams[:, 1, 0, :3, 3] += torch.tensor([400.0, -250.0, 0.0])
        hyp = hyp * (1.0 + 0.3 * torch.rand(hyp.shape, generator=S._gen(5)))
    net = StageNet(dict(STAGE_ARGS), S.NDEPTHS[s], s).eval()
    net.load_state_dict(S.fill_state_dict(net.state_dict(), seed=3))
    net = net.to(DEV)
    res = {}
    for layout in ("nchw", "cl"):
        config.set_cv_layout(layout)
        try:
            res[layout] = net.build_cost_volume(feats.to(DEV), cams.to(DEV), hyp.to(DEV))
        finally:
            config.set_cv_layout("cl")
    (v0, s0, e0, w0), (v1, s1, e1, w1) = res["nchw"], res["cl"]
    assert rel_l1(e1, e0) < 1e-5, rel_l1(e1, e0)
    assert rel_l1(w1, w0) < 1e-5
    assert rel_l1(v1, v0) < 2e-5, rel_l1(v1, v0)
    # the channels-last kernels take the argmax of the similarity inside the kernel: compare the chosen hypotheses
    from mvsformer_b200 import engine
    sd0 = engine.argmax_gather(s0, hyp.to(DEV))
    assert s1.shape == sd0.shape
    assert (s1 == sd0).float().mean() > (0.97 if wild else 0.99)          # exact ties flip under another summation order
    assert torch.isfinite(v1).all()


# ---- fused visibility net (csrc/vis_fused.cu) ----------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(4, 144, 192), (3, 37, 50), (2, 30, 14), (1, 31, 15), (5, 200, 333)])
def test_fused_vis_net_matches_the_fp32_kernel(shape):
    """The fused tcgen05 kernel (TF32 operands for the 16->16 and 16->8 layers, fp32 accumulation) vs the FP32 CUDA-core
    kernel of csrc/vis_net.cu (which tests/test_gpu_parity.py holds to the oracle at 1e-5): the difference is the TF32
    rounding of two layers' operands (2^-11 relative each) seen through the sigmoid.  Incl. partial tiles and maps smaller
    than a tile."""
    from tests.helpers import rel_l1
    m, h, w = shape
    net = StageNet(dict(STAGE_ARGS), 8, 2).eval()
    net.load_state_dict(S.fill_state_dict(net.state_dict(), seed=21))
    net = net.to(DEV)
    ent = (2.0 * torch.rand(1, m, h, w, generator=S._gen(h + w))).to(DEV)
    old = config.conv_precision()
    config.set_conv_precision("tf32")
    try:
        config.set_vis_fused(False)
        want = net._vis_weight(ent)
        config.set_vis_fused(True)
        got = net._vis_weight(ent)
    finally:
        config.set_vis_fused(True)
        config.set_conv_precision(old)
    torch.cuda.synchronize()
    assert got.shape == want.shape
    # (against the unfused TF32 route it replaced, same operands, the fused kernel agreed to 2e-5: profiles/r02_gpu_tests.log)
    assert rel_l1(got, want) < 3e-3, rel_l1(got, want)
    assert float((got - want).abs().max()) < 3e-2
