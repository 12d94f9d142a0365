"""The drop-in boundary end to end, on the CPU: the UNMODIFIED reference ``TwinMVSNet`` (FPN + Twins
backbone + cascade loop, models/mvsformer_model.py:310-449) is built twice from the shipped config — once as is, once
after ``mvsformer_b200.mvsformer_model.install_into`` rebinds ``StageNet`` and the schedulers — with the same weights
(``load_state_dict(strict=True)`` across the two).  A training step (forward, cross-entropy on every stage's
``prob_volume_pre``, backward) must give the same stage outputs and the same gradients in the reference's own backbone,
i.e. our kernels (run here on the CPU thread emulation of their source) slot into the reference's autograd graph.
Skipped where /root/reference is absent (the GPU box)."""
import json
import os

import pytest
import torch
import torch.nn.functional as F

from mvsformer_b200 import synthetic as S
from mvsformer_b200.mvsformer_model import StageNet, install_into
from oracle import ref_import
from tests.emu import harness
from tests.helpers import rel_l1

pytestmark = pytest.mark.skipif(not ref_import.reference_available(), reason="reference not present (build container only)")


def _inputs(height, width, views):
    g = S._gen(77)
    imgs = torch.rand(1, views, 3, height, width, generator=g)
    cams = S.make_cameras(1, views, height, width)
    return imgs, cams, S.make_depth_range(1)


def _step(model, imgs, cams, dv, targets):
    torch.manual_seed(5)                                        # any stochastic layer of the backbone draws the same numbers
    out = model(imgs, cams, dv)
    loss = sum(F.cross_entropy(out["stage%d" % (s + 1)]["prob_volume_pre"], targets[s]) for s in range(4))
    model.zero_grad(set_to_none=True)
    loss.backward()
    return out, float(loss.detach())


def test_install_into_reference_twinmvsnet_training_step(monkeypatch):
    ns = ref_import.load_reference()
    R = ns.mvsformer_model
    args = json.load(open(os.path.join(ref_import.REFERENCE_ROOT, "configs", "config_mvsformer.json")))["arch"]["args"]
    height, width, views = 128, 192, 3
    imgs, cams, dv = _inputs(height, width, views)
    targets = [torch.randint(0, S.NDEPTHS[s], (1,) + S.stage_hw(height, width, s), generator=S._gen(s)) for s in range(4)]

    torch.manual_seed(1)
    ref_model = R.TwinMVSNet(args).train()
    want, want_loss = _step(ref_model, imgs, cams, dv, targets)
    want_grads = {k: p.grad.clone() for k, p in ref_model.named_parameters() if p.grad is not None}

    saved = {n: getattr(R, n) for n in ("StageNet", "init_inverse_range", "init_range", "schedule_inverse_range", "schedule_range",
                                         "CostRegNet", "CostRegNet3D", "CostRegNet2D")}
    try:
        install_into(R)
        harness.install(monkeypatch.setattr)                    # kernels -> CPU emulation of the same source
        torch.manual_seed(1)
        ours = R.TwinMVSNet(args).train()
        assert all(isinstance(f, StageNet) for f in ours.fusions)
        ours.load_state_dict(ref_model.state_dict(), strict=True)            # reference checkpoint loads unchanged
        got, got_loss = _step(ours, imgs, cams, dv, targets)
    finally:
        for n, v in saved.items():
            setattr(R, n, v)

    assert set(got.keys()) == set(want.keys())
    # stage 1 sees identical hypotheses: strict; later stages inherit argmax depths (discontinuous), compare robustly
    assert rel_l1(got["stage1"]["prob_volume_pre"], want["stage1"]["prob_volume_pre"]) < 1e-4
    assert (got["stage1"]["depth"] == want["stage1"]["depth"]).float().mean() > 0.99
    for s in range(1, 4):
        a, b = got["stage%d" % (s + 1)]["prob_volume_pre"], want["stage%d" % (s + 1)]["prob_volume_pre"]
        close = ((a - b).abs() <= 1e-3 * b.abs().mean()).float().mean()
        assert close > 0.9, (s, float(close))
    assert got_loss == pytest.approx(want_loss, rel=2e-3)
    # gradients arrive in the reference's own FPN encoder and Twins backbone through our backward kernels
    checked = 0
    for name, p in ours.named_parameters():
        if name.startswith("fusions.") or name not in want_grads or p.grad is None:
            continue
        if want_grads[name].abs().mean() < 1e-7:               # e.g. a conv bias in front of a BatchNorm: exactly zero, rounding noise
            continue
        assert rel_l1(p.grad, want_grads[name]) < 5e-2, name
        checked += 1
    assert checked > 20
    # and in our own modules' parameters, against the reference modules' gradients
    # (the visibility net's gradients are heavily cancelling sums, DESIGN.md §4.4: looser bar)
    for name, tol in (("fusions.0.cost_reg.conv1.conv.weight", 5e-3), ("fusions.0.cost_reg.prob.weight", 5e-3),
                      ("fusions.0.vis.0.conv.weight", 2e-2)):
        assert rel_l1(dict(ours.named_parameters())[name].grad, want_grads[name]) < tol, name
