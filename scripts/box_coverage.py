"""Fast-path coverage of the cost-volume kernel's staged box on the bench workload (GPU).

For every stage of one cascade pass at the bench configuration, recompute the sample positions of
all (view, hypothesis, pixel) triples and report, for candidate box sizes BW x BH, the fraction whose
2x2 footprint lies inside the CTA's box (origin = tile minimum, x rounded down to a multiple of 4,
as the cost-volume kernels do).  Samples outside take the predicated global path."""
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

DG = (8, 4, 2, 1)
CAND = {0: [(112, 8), (128, 8), (160, 8)], 1: [(64, 8), (96, 8), (64, 12)], 2: [(48, 12), (64, 12), (96, 12), (64, 16)],
        3: [(48, 16), (64, 16), (96, 16), (64, 24), (96, 24), (128, 24)]}
for s in range(4):
    hyp = out["stage%d" % (s + 1)]["depth_values"][0]                       # [D, h, w]
    d, h, w = hyp.shape
    rel = engine.relative_projections(cams["stage%d" % (s + 1)])[0]          # [N, 12]
    ys, xs = torch.meshgrid(torch.arange(h, device=dev, dtype=torch.float32),
                            torch.arange(w, device=dev, dtype=torch.float32), indexing="ij")
    th = 8 // DG[s]
    print("stage %d: D=%d %dx%d, tile 32x%d, depth std within 32x%d tiles (mm): %.2f" % (
        s + 1, d, h, w, th, th, float(out["stage%d" % (s + 1)]["depth"][0].unfold(0, th, th).unfold(1, 32, 32).std(dim=(-1, -2)).mean())))
    for v in range(rel.shape[0]):
        m = rel[v]
        rx = m[0] * xs + m[1] * ys + m[2]
        ry = m[4] * xs + m[5] * ys + m[6]
        rz = m[8] * xs + m[9] * ys + m[10]
        den = rz[None] * hyp + m[11] + 1e-6
        ix = (rx[None] * hyp + m[3]) / den                                  # pixel units (align_corners round trip is the identity)
        iy = (ry[None] * hyp + m[7]) / den
        x0, y0 = ix.floor(), iy.floor()
        # tile minima over (all D, 32 x th pixels)
        def tile_min(t):
            t = t.amin(dim=0)
            hp, wp = (h + th - 1) // th * th, (w + 31) // 32 * 32
            t = torch.nn.functional.pad(t, (0, wp - w, 0, hp - h), value=float("inf"))
            t = t.view(hp // th, th, wp // 32, 32).amin(dim=(1, 3))
            return t.repeat_interleave(th, 0).repeat_interleave(32, 1)[:h, :w]
        bx = torch.floor(tile_min(x0) / 4) * 4
        by = tile_min(y0)
        lx, ly = x0 - bx[None], y0 - by[None]
        line = []
        for bw, bh in CAND[s]:
            ok = (lx + 1 < bw) & (ly + 1 < bh)
            line.append("%dx%d: %.4f" % (bw, bh, float(ok.float().mean())))
        print("   view %d  x-span p50/p99 %.0f/%.0f  y-span p50/p99 %.0f/%.0f   %s" % (
            v + 1, float(lx.amax(0).flatten().quantile(0.5)), float(lx.amax(0).flatten()[::7].quantile(0.99)),
            float(ly.amax(0).flatten().quantile(0.5)), float(ly.amax(0).flatten()[::7].quantile(0.99)), "  ".join(line)))
