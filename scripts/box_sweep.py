"""Times the cost-volume kernels (pass A, pass B) per stage for every box variant selectable with
MVS_K1_BOX, on the hypotheses of one bench-configuration cascade pass (GPU)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from mvsformer_b200 import config, engine  # noqa: E402
from mvsformer_b200 import synthetic as S  # noqa: E402

config.set_conv_precision("tf32")
dev = torch.device("cuda", 0)
net = bench.build_engine(dev)
feats, cams, dv = bench.host_inputs(bench.HEIGHT, bench.WIDTH, bench.VIEWS, 1234, pin=False)
feats = {k: v.to(dev) for k, v in feats.items()}
cams = {k: v.to(dev) for k, v in cams.items()}
with torch.no_grad():
    out = net(feats, cams, dv.to(dev), tmp=list(S.EVAL_TMP))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
NVAR = (4, 2, 4, 4)


def timed(fn, n=12):
    ts = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); r = fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2], r


for s in range(4):
    f = feats["stage%d" % (s + 1)]
    hyp = out["stage%d" % (s + 1)]["depth_values"]
    rel = engine.relative_projections(cams["stage%d" % (s + 1)])
    base = None
    for var in range(NVAR[s]):
        sel = [0, 0, 0, 0]
        sel[s] = var
        os.environ["MVS_K1_BOX"] = ",".join(map(str, sel))
        ta, (ent, sim) = timed(lambda: engine.cost_volume_entropy(f, rel, hyp, 8, True))
        wgt = net.fusions[s]._vis_weight(ent)
        tb, vol = timed(lambda: engine.cost_volume_aggregate(f, rel, hyp, wgt, 8, round_tf32=True))
        if base is None:
            base = (ent, sim, vol)
        same = all(torch.equal(x, y) for x, y in zip((ent, sim, vol), base))
        print("stage %d variant %d: passA %.3f ms  passB %.3f ms  sum %.3f  identical_to_v0=%s" % (s + 1, var, ta, tb, ta + tb, same))
