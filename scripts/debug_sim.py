"""Debug: stage-1 cosine-similarity volume of the channels-last kernels vs the NCHW kernels on the golden case."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mvsformer_b200 import config, synthetic as S
from mvsformer_b200.mvsformer_model import StageNet
from tests.helpers import STAGE_ARGS, load_golden
for s in range(4):
    g = load_golden("stage%d.npz" % (s + 1))
    batch = int(g["batch"])
    feats = S.make_features(batch, 3, 128, 192, stages=(s,))["stage%d" % (s + 1)].cuda()
    cams = S.make_cameras(batch, 3, 128, 192)["stage%d" % (s + 1)].cuda()
    hyp = S.narrow_hypotheses(s, 128, 192, batch).cuda()
    net = StageNet(dict(STAGE_ARGS), S.NDEPTHS[s], s).eval()
    net.load_state_dict(S.fill_state_dict(net.state_dict(), seed=s), strict=True)
    net = net.cuda()
    res = {}
    for layout in ("nchw", "cl"):
        config.set_cv_layout(layout)
        v, sim, e, w = net.build_cost_volume(feats, cams, hyp)
        out = net(feats, cams, hyp, tmp=list(S.EVAL_TMP))
        agree = (out["sim_depth"].cpu() == torch.from_numpy(g["sim_depth"])).float().mean().item()
        res[layout] = (v, sim, e)
        print("stage", s + 1, layout, "sim_depth agreement with golden", agree, "batch", batch, "shape", tuple(sim.shape))
    d = (res["cl"][1] - res["nchw"][1]).abs()
    print("   sim max abs diff", d.max().item(), "at", (d == d.max()).nonzero()[0].tolist(), "sim abs mean", res["nchw"][1].abs().mean().item(),
          "| volume max abs diff", (res["cl"][0] - res["nchw"][0]).abs().max().item(), "| entropy max abs diff", (res["cl"][2] - res["nchw"][2]).abs().max().item())
    top2 = res["nchw"][1].topk(2, dim=1)[0]
    print("   smallest top1-top2 gap (nchw):", (top2[:, 0] - top2[:, 1]).min().item())
