"""BASELINE.json configs 2, 3 and 4 at FULL size on the device, compared with the CPU oracle (not with invariants):
the oracle port finishes one full-size cascade in seconds on the GPU box's host cores.  Both conv precisions the
package ships: the fp32-grade default (tf32x3) and the TF32 mode bench.py times (north-star bar: refined depth within
1e-3 relative L1).  Also: the unmodified reference modules (oracle/_ref, when staged) on the SAME GPU."""
import pytest
import torch

from mvsformer_b200 import config, synthetic as S
from mvsformer_b200.mvsformer_model import CascadeMVS
from tests.helpers import CASCADE_ARGS, rel_l1

pytestmark = pytest.mark.gpu
DEV = "cuda"

CONFIGS = [("cfg2 DTU", 1, 5, 1152, 1536), ("cfg3 BlendedMVS", 4, 7, 576, 768), ("cfg4 T&T @1088", 1, 11, 1088, 1920)]


def _net_and_sds():
    net = CascadeMVS(dict(CASCADE_ARGS)).eval()
    sds, full = [], {}
    for s in range(4):
        sd = S.fill_state_dict(net.fusions[s].state_dict(), seed=s)
        sds.append(sd)
        full.update({"fusions.%d.%s" % (s, k): v for k, v in sd.items()})
    net.load_state_dict(full, strict=True)
    return net.to(DEV), sds


@pytest.mark.parametrize("name,batch,views,height,width", CONFIGS)
def test_full_size_cascade_vs_oracle(name, batch, views, height, width):
    from oracle import mvs_oracle as O

    net, sds = _net_and_sds()
    feats = S.make_features(batch, views, height, width, seed=1234)
    cams = S.make_cameras(batch, views, height, width)
    dv = S.make_depth_range(batch)
    with torch.no_grad():
        want = O.cascade_forward(feats, cams, dv, sds)
    fd = {k: v.to(DEV) for k, v in feats.items()}
    cd = {k: v.to(DEV) for k, v in cams.items()}
    saved = config.conv_precision()
    try:
        for mode, tol_depth, tol_conf in (("tf32x3", 1e-5, 1e-4), ("tf32", 1e-3, 5e-3)):
            config.set_conv_precision(mode)
            with torch.no_grad():
                out = net(fd, cd, dv.to(DEV), tmp=list(S.EVAL_TMP))
            torch.cuda.synchronize()
            rd = rel_l1(out["refined_depth"].cpu(), want["refined_depth"])
            rc = rel_l1(out["photometric_confidence"].cpu(), want["photometric_confidence"])
            print("%s %s: refined depth rel-L1 %.3e, confidence rel-L1 %.3e" % (name, mode, rd, rc))
            assert rd < tol_depth, (name, mode, rd)
            assert rc < tol_conf, (name, mode, rc)
            for s in range(4):
                st = out["stage%d" % (s + 1)]
                assert rel_l1(st["depth"].cpu(), want["stage%d" % (s + 1)]["depth"]) < tol_depth * 2
    finally:
        config.set_conv_precision(saved)


def test_cfg2_vs_unmodified_reference_on_same_gpu():
    """The staged, unmodified reference modules run on cuda:0 (PyTorch eager + cuDNN, TF32 convs as torch defaults)
    against the engine in its TF32 mode: both are within the north-star tolerance of each other."""
    from oracle import ref_cascade as R

    if not R.available():
        pytest.skip("oracle/_ref not staged")
    net, sds = _net_and_sds()
    height, width, batch, views = 1152, 1536, 1, 5
    fd = {k: v.to(DEV) for k, v in S.make_features(batch, views, height, width, seed=1234).items()}
    cd = {k: v.to(DEV) for k, v in S.make_cameras(batch, views, height, width).items()}
    dv = S.make_depth_range(batch).to(DEV)
    nets = R.build_stage_nets(sds, DEV)
    with torch.no_grad():
        want = R.cascade(nets, fd, cd, dv)
    saved = config.conv_precision()
    try:
        config.set_conv_precision("tf32")
        with torch.no_grad():
            out = net(fd, cd, dv, tmp=list(S.EVAL_TMP))
        torch.cuda.synchronize()
    finally:
        config.set_conv_precision(saved)
    rd = rel_l1(out["refined_depth"], want["refined_depth"])
    print("cfg2 engine(tf32) vs reference-on-GPU(cuDNN TF32): refined depth rel-L1 %.3e" % rd)
    assert rd < 2e-3          # two independently TF32-rounded paths, each within 1e-3 of fp32
