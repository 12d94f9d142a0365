"""GPU tests of opt-in code paths that have been compiled and reviewed but NOT yet executed on a B200
(written after the round's GPU minutes were spent).  They only run with MVS_TEST_EXPERIMENTAL=1, so that an
unverified kernel can never take the verified suite down with it; once a path has passed here on the device its
test moves into the regular files.

* MVS_CV_STORE: cost-volume build with one sampling pass (pass A stores the per-view correlation, the
  aggregation streams over it) — must be BIT-identical to the two-pass build (same FMA sequence).
"""
import os

import pytest
import torch

from mvsformer_b200 import config, synthetic as S
from mvsformer_b200.mvsformer_model import StageNet
from tests.helpers import STAGE_ARGS

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("MVS_TEST_EXPERIMENTAL", "0") in ("", "0"),
                                 reason="opt-in path not yet run on a GPU; set MVS_TEST_EXPERIMENTAL=1")]
DEV = "cuda"


def _build(net, feats, cams, hyp, store):
    config.set_cv_store(store)
    try:
        return net.build_cost_volume(feats.to(DEV), cams.to(DEV), hyp.to(DEV))
    finally:
        config.set_cv_store(False)


@pytest.mark.parametrize("s", [0, 1, 2])
@pytest.mark.parametrize("mode", ["tf32x3", "tf32"])
@pytest.mark.parametrize("wild", [False, True])
def test_cv_store_is_bit_identical_to_two_pass(s, mode, wild):
    height, width, batch, views = 128, 192, 2, 4
    feats = S.make_features(batch, views, height, width, stages=(s,), smooth=not wild)["stage%d" % (s + 1)]
    cams = S.make_cameras(batch, views, height, width)["stage%d" % (s + 1)].clone()
    hyp = S.narrow_hypotheses(s, height, width, batch)
    if wild:                                    # samples outside the staged box / the image: predicated global path
        cams[:, 2, 0, 0, 3] += 500.0
        cams[:, 3, 0, 2, 3] -= 900.0
    net = StageNet(dict(STAGE_ARGS), S.NDEPTHS[s], s).eval()
    net.load_state_dict(S.fill_state_dict(net.state_dict(), seed=s))
    net = net.to(DEV)
    old = config.conv_precision()
    config.set_conv_precision(mode)
    try:
        vol2, sim2, ent2, w2 = _build(net, feats, cams, hyp, False)
        vol1, sim1, ent1, w1 = _build(net, feats, cams, hyp, True)
    finally:
        config.set_conv_precision(old)
    assert torch.equal(ent1, ent2) and torch.equal(sim1, sim2) and torch.equal(w1, w2)
    assert torch.equal(vol1, vol2)


def test_cv_store_falls_back_at_stage4():
    """C/G = 1 (stage 4): the correlation would be as large as the warped tensor, so the two-pass build runs."""
    s, height, width = 3, 64, 96
    feats = S.make_features(1, 3, height, width, stages=(s,))["stage4"]
    cams = S.make_cameras(1, 3, height, width)["stage4"]
    hyp = S.narrow_hypotheses(s, height, width, 1)
    net = StageNet(dict(STAGE_ARGS), 4, s).eval().to(DEV)
    a = _build(net, feats, cams, hyp, True)[0]
    b = _build(net, feats, cams, hyp, False)[0]
    assert torch.equal(a, b)
