"""Pin the CPU oracle against golden vectors produced by the unmodified reference
(oracle/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from mvsformer_b200 import synthetic as S
from oracle import mvs_oracle as O
from tests.helpers import checksum, load_golden, max_abs, rel_l1


def _compose(pair):
    return O.compose_projection(pair)


def test_warp_golden():
    g = load_golden("warp.npz")
    feats = S.make_features(2, 3, 24, 40, seed=5, stages=(3,), feat_chs=(0, 0, 0, 8))["stage4"]
    assert checksum(feats) == pytest.approx(float(g["feat_checksum"]), rel=1e-12)
    cams = S.make_cameras(2, 3, 24, 40)["stage4"].clone()
    cams[:, 2, 0, 0, 3] += 150.0
    ref_p = _compose(cams[:, 0])
    dv_map = S.make_depth_range(2)[:, ::32][:, :6]
    dv_px = dv_map.view(2, 6, 1, 1) * (1.0 + 0.1 * torch.rand(2, 6, 24, 40, generator=S._gen(3)))
    for v in (1, 2):
        src_p = _compose(cams[:, v])
        for tag, dv in (("bd", dv_map), ("px", dv_px)):
            warped, mask = O.homo_warping_3D_with_mask(feats[:, v], src_p, ref_p, dv)
            gw = torch.from_numpy(g["warped_%s_v%d" % (tag, v)])
            gm = torch.from_numpy(g["mask_%s_v%d" % (tag, v)])
            assert rel_l1(warped, gw) < 2e-6
            assert max_abs(warped, gw) < 2e-4
            assert (mask != gm).float().mean() < 1e-3          # pixels within 1 ulp of the border may flip


def test_schedules_golden():
    g = load_golden("schedules.npz")
    dv = S.make_depth_range(2)
    assert max_abs(O.init_inverse_range(dv, 32, 8, 12), g["init_inverse_range"]) < 2e-4
    assert max_abs(O.init_range(dv, 32, 8, 12), g["init_range"]) < 2e-4
    hyp = torch.from_numpy(g["init_inverse_range"])
    depth = torch.from_numpy(g["sched_depth_in"])
    assert rel_l1(O.schedule_inverse_range(depth, hyp, 16, 2.67, 16, 24), g["schedule_inverse_range"]) < 1e-6
    assert rel_l1(O.schedule_range(depth, 16, 2.67 * (dv[:, 1] - dv[:, 0]), 16, 24), g["schedule_range"]) < 1e-6
    p = torch.from_numpy(g["prob_in"])
    assert rel_l1(O.depth_regression(p, hyp), g["depth_regression_map"]) < 1e-6
    assert rel_l1(O.depth_regression(p, dv[:, :32]), g["depth_regression_vec"]) < 1e-6
    for n in (2, 3, 4):
        assert rel_l1(O.conf_regression(p, n), g["conf_regression_n%d" % n]) < 1e-6


def _stage_case(s, batch, train=False):
    from tests.helpers import reference_state_dict_template

    height, width = 128, 192
    feats = S.make_features(batch, 3, height, width, stages=(s,))["stage%d" % (s + 1)]
    cams = S.make_cameras(batch, 3, height, width)["stage%d" % (s + 1)]
    hyp = S.narrow_hypotheses(s, height, width, batch)
    sd = S.fill_state_dict(reference_state_dict_template(s, S.NDEPTHS[s]), seed=s)
    return feats, cams, hyp, sd


@pytest.mark.parametrize("s", [0, 1, 2, 3])
def test_stage_golden(s):
    g = load_golden("stage%d.npz" % (s + 1))
    feats, cams, hyp, sd = _stage_case(s, int(g["batch"]))
    assert checksum(feats) == pytest.approx(float(g["feat_checksum"]), rel=1e-12)
    out = O.stage_forward(feats, cams, hyp, sd, S.NDEPTHS[s], S.EVAL_TMP[s])
    assert rel_l1(out["prob_volume_pre"], g["prob_volume_pre"]) < 1e-5
    assert rel_l1(out["prob_volume"], g["prob_volume"]) < 1e-5
    assert rel_l1(out["depth"], g["depth"]) < 1e-6
    assert rel_l1(out["photometric_confidence"], g["photometric_confidence"]) < 1e-5
    agree = (out["sim_depth"] == torch.from_numpy(g["sim_depth"])).float().mean()
    assert agree > 0.999


def test_stage_train_golden():
    g = load_golden("stage2_train.npz")
    feats, cams, hyp, sd = _stage_case(1, 2)
    out = O.stage_forward(feats, cams, hyp, sd, S.NDEPTHS[1], S.EVAL_TMP[1], training=True)
    assert rel_l1(out["prob_volume_pre"], g["prob_volume_pre"]) < 1e-5
    agree = (out["depth"] == torch.from_numpy(g["depth"])).float().mean()      # argmax depth
    assert agree > 0.999


def test_cascade_golden():
    from tests.helpers import reference_state_dict_template

    g = load_golden("cascade.npz")
    height, width, batch, views = int(g["height"]), int(g["width"]), int(g["batch"]), int(g["views"])
    feats = S.make_features(batch, views, height, width, seed=int(g["feat_seed"]))
    assert checksum(feats["stage4"]) == pytest.approx(float(g["feat_checksum"]), rel=1e-12)
    cams = S.make_cameras(batch, views, height, width)
    dv = S.make_depth_range(batch)
    sds = [S.fill_state_dict(reference_state_dict_template(s, S.NDEPTHS[s]), seed=int(g["weight_seed0"]) + s)
           for s in range(4)]
    out = O.cascade_forward(feats, cams, dv, sds)
    for s in range(4):
        assert rel_l1(out["stage%d" % (s + 1)]["depth_values"], g["stage%d_depth_values" % (s + 1)]) < 1e-6
        assert rel_l1(out["stage%d" % (s + 1)]["depth"], g["stage%d_depth" % (s + 1)]) < 1e-6
    assert rel_l1(out["refined_depth"], g["refined_depth"]) < 1e-6
    assert rel_l1(out["photometric_confidence"], g["photometric_confidence"]) < 1e-5


def _train_grad_case(s):
    """Inputs of oracle/make_golden.py::gen_train_grads, regenerated from the recorded seeds."""
    from tests.helpers import reference_state_dict_template

    g = load_golden("stage%d_train_grads.npz" % (s + 1))
    height, width, batch = int(g["height"]), int(g["width"]), int(g["batch"])
    feats = S.make_features(batch, 3, height, width, seed=int(g["feat_seed"]), stages=(s,))["stage%d" % (s + 1)]
    assert checksum(feats) == pytest.approx(float(g["feat_checksum"]), rel=1e-12)
    cams = S.make_cameras(batch, 3, height, width)["stage%d" % (s + 1)]
    hyp = S.narrow_hypotheses(s, height, width, batch)
    sd = S.fill_state_dict(reference_state_dict_template(s, S.NDEPTHS[s]), seed=int(g["weight_seed"]))
    target = torch.randint(0, S.NDEPTHS[s], (batch, feats.shape[-2], feats.shape[-1]), generator=S._gen(700 + s))
    return g, feats, cams, hyp, sd, target


@pytest.mark.parametrize("s", [1, 3])
def test_training_gradients_golden(s):
    """torch autograd over the oracle restatement == torch autograd over the unmodified reference
    (train mode, CE loss on prob_volume_pre): pins the gradient oracle used by the training-path tests."""
    g, feats, cams, hyp, sd, target = _train_grad_case(s)
    feats = feats.clone().requires_grad_(True)
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()
              if v.dtype.is_floating_point and "running" not in k}
    sd2 = dict(sd)
    sd2.update(params)
    out = O.stage_forward(feats, cams, hyp, sd2, S.NDEPTHS[s], S.EVAL_TMP[s], training=True)
    assert rel_l1(out["prob_volume_pre"], g["prob_volume_pre"]) < 1e-5
    loss = torch.nn.functional.cross_entropy(out["prob_volume_pre"], target)
    assert float(loss.detach()) == pytest.approx(float(g["loss"]), rel=1e-5)
    loss.backward()
    assert rel_l1(feats.grad, g["grad_features"]) < 1e-4
    for name in [str(n) for n in g["param_names"]]:
        if name == "cost_reg.prob.bias":
            continue                                        # exactly zero in exact arithmetic (softmax shift invariance)
        want_abs = float(g["abs_sum/" + name])
        got_abs = float(params[name].grad.double().abs().sum())
        # the visibility net's gradients are ~1e-6 sums with heavy cancellation: 1e-2; everything else 2e-3
        tol = 1e-2 if name.startswith("vis.") else 2e-3
        assert got_abs == pytest.approx(want_abs, rel=tol), name
        if "grad/" + name in g.files:
            assert rel_l1(params[name].grad, g["grad/" + name]) < tol, name
