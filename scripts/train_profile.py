"""Per-entry-point CUDA-event breakdown of one cfg-5-shaped training step (B=1, 5 views, 512x640, 4-stage
cascade, CE loss, SGD).  Run on the GPU box:  python scripts/train_profile.py > gpurun_out/train_profile.json"""
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mvsformer_b200 import autograd, synthetic as S  # noqa: E402
from mvsformer_b200.mvsformer_model import CascadeMVS  # noqa: E402

ARGS = {"base_ch": 8, "fusion_type": "cnn", "depth_type": "ce", "ndepths": list(S.NDEPTHS),
        "depth_interals_ratio": list(S.DEPTH_INTERVAL_RATIO), "inverse_depth": True}


def main():
    dev = torch.device("cuda", 0)
    height, width, views = 512, 640, 5
    feats = {k: v.to(dev).requires_grad_(True) for k, v in S.make_features(1, views, height, width, seed=5).items()}
    cams = {k: v.to(dev) for k, v in S.make_cameras(1, views, height, width).items()}
    dv = S.make_depth_range(1).to(dev)
    net = CascadeMVS(dict(ARGS)).train().to(dev)
    targets = [torch.randint(0, S.NDEPTHS[s], (1,) + S.stage_hw(height, width, s), generator=S._gen(s)).to(dev)
               for s in range(4)]

    def step():
        net.zero_grad(set_to_none=True)
        out = net(feats, cams, dv)
        sum(F.cross_entropy(out["stage%d" % (s + 1)]["prob_volume_pre"], targets[s]) for s in range(4)).backward()

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    step()
    e1.record()
    torch.cuda.synchronize()
    with autograd.profile_kernels() as prof:
        step()
    summ = prof.summary()
    total = sum(ms for ms, _ in summ.values())
    rows = {k: {"ms": round(ms, 4), "launches": n, "share": round(ms / total, 4)} for k, (ms, n) in
            sorted(summ.items(), key=lambda kv: -kv[1][0])}
    print(json.dumps({"step_ms_unprofiled": e0.elapsed_time(e1), "kernel_ms_sum": total, "entry_points": rows}, indent=1))


if __name__ == "__main__":
    main()
