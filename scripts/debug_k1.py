"""Debug helper: run one stage's cost-volume build with the TMA kernels and with the generic
kernels and print the differences (run under compute-sanitizer when hunting a fault)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mvsformer_b200 import synthetic as S  # noqa: E402
from mvsformer_b200.mvsformer_model import StageNet  # noqa: E402

s = int(sys.argv[1]) if len(sys.argv) > 1 else 3
height, width = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (128, 192)
feats = S.make_features(1, 3, height, width, stages=(s,))["stage%d" % (s + 1)].cuda()
cams = S.make_cameras(1, 3, height, width)["stage%d" % (s + 1)].cuda()
hyp = S.narrow_hypotheses(s, height, width, 1).cuda()
net = StageNet({"base_ch": 8, "fusion_type": "cnn", "depth_type": "ce"}, S.NDEPTHS[s], s).eval()
net.load_state_dict(S.fill_state_dict(net.state_dict(), seed=s))
net = net.cuda()
os.environ["MVS_K1_IMPL"] = "generic"
ref = net.build_cost_volume(feats, cams, hyp)
torch.cuda.synchronize()
print("generic ok")
os.environ["MVS_K1_IMPL"] = "tma"
got = net.build_cost_volume(feats, cams, hyp)
torch.cuda.synchronize()
print("tma ok")
for name, a, b in zip(("volume", "sim", "entropy", "weight"), got, ref):
    print(name, float((a - b).abs().max()), float((a - b).abs().mean() / b.abs().mean()))
