"""Per-layer timing of the depth-unstrided convolutions at the cfg-2 shapes: the generic tcgen05 kernels (conv3d_tc.cu) vs the persistent TMA-fed kernels (conv3d_tma.cu).

    python scripts/conv_bench.py [out.json]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from mvsformer_b200 import engine

DEV = "cuda"
# name, cin, cout, kd, B, D, H, W   (stage 4: D=4 at 1152x1536; stage 3: D=8 at 576x768; visibility net: 4 maps as batch)
CASES = [("s4.conv2", 16, 16, 3, 1, 4, 576, 768), ("s4.conv4", 32, 32, 3, 1, 4, 288, 384), ("s4.conv6", 64, 64, 3, 1, 4, 144, 192),
         ("s3.conv2", 16, 16, 3, 1, 8, 288, 384), ("s3.conv4", 32, 32, 3, 1, 8, 144, 192), ("s3.conv6", 64, 64, 3, 1, 8, 72, 96),
         ("s4.vis2", 16, 16, 1, 4, 1, 1152, 1536), ("s4.vis3", 16, 8, 1, 4, 1, 1152, 1536), ("s3.vis2", 16, 16, 1, 4, 1, 576, 768)]


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


# stride-2 and transposed layers: name, cin, cout, B, D, H, W (input size)
S2_CASES = [("s4.conv1", 8, 16, 1, 4, 1152, 1536), ("s4.conv3", 16, 32, 1, 4, 576, 768), ("s4.conv5", 32, 64, 1, 4, 288, 384),
            ("s3.conv1", 8, 16, 1, 8, 576, 768), ("s3.conv3", 16, 32, 1, 8, 288, 384), ("s3.conv5", 32, 64, 1, 8, 144, 192)]
DE_CASES = [("s4.deconv7", 64, 32, 1, 4, 144, 192), ("s4.deconv9", 32, 16, 1, 4, 288, 384), ("s4.deconv11", 16, 8, 1, 4, 576, 768),
            ("s3.deconv7", 64, 32, 1, 8, 72, 96), ("s3.deconv9", 32, 16, 1, 8, 144, 192), ("s3.deconv11", 16, 8, 1, 8, 288, 384)]


def strided_cases(out):
    for cases, mode in ((S2_CASES, engine.TMA_S2), (DE_CASES, engine.TMA_DECONV)):
        for name, cin, cout, b, d, h, w in cases:
            g = torch.Generator().manual_seed(1)
            wt = engine.round_tf32(torch.randn(3, 3, 3, cin, cout, generator=g) * 0.05).to(DEV)
            x = engine.round_tf32(torch.randn(b, d, h, w, cin, generator=g)).to(DEV)
            shift = torch.zeros(cout, device=DEV)
            wn, nt = engine.pack_tma_weights(wt, mode)
            y_new = engine.conv3d_tma(x, wn, nt, cout, 3, shift, None, True, mode)
            skip = torch.randn_like(y_new) if mode == engine.TMA_DECONV else None       # the up path adds a skip tensor
            y_new = engine.conv3d_tma(x, wn, nt, cout, 3, shift, skip, True, mode)
            res = {"tma_us": timed(lambda: engine.conv3d_tma(x, wn, nt, cout, 3, shift, skip, True, mode))}
            if mode == engine.TMA_S2:
                hi, _, ntk = engine.pack_tc_weights(wt, False)
                old = lambda: engine.conv3d_tc(x, hi, None, ntk, cout, 3, shift, None, (1, 2, 2), True)
            else:
                hi, _, ntk = engine.pack_tc_deconv_weights(wt, False)
                old = lambda: engine.deconv3d_tc(x, hi, None, ntk, cout, 3, shift, skip, 1, True)
            y_old = old()
            res["old_us"] = timed(old)
            res["max_abs_diff"] = float((y_new.reshape(-1) - y_old.reshape(-1)).abs().max())
            res["tma_gbs"] = 4.0 * (x.numel() + y_new.numel()) / res["tma_us"] / 1e3
            out[name] = res
            print(name, {k: (round(v, 2) if isinstance(v, float) else v) for k, v in res.items()}, flush=True)


def main():
    out = {}
    if "--strided-only" not in sys.argv:
        unstrided_cases(out)
    strided_cases(out)
    outs = [a for a in sys.argv[1:] if not a.startswith("--")]
    if outs:
        json.dump(out, open(outs[0], "w"), indent=1)


def unstrided_cases(out):
    only = [a.split("=", 1)[1] for a in sys.argv if a.startswith("--only=")]
    for name, cin, cout, kd, b, d, h, w in CASES:
        if only and name not in only:
            continue
        g = torch.Generator().manual_seed(1)
        wt = engine.round_tf32(torch.randn(kd, 3, 3, cin, cout, generator=g) * 0.05).to(DEV)
        x = engine.round_tf32(torch.randn(b, d, h, w, cin, generator=g)).to(DEV)
        shift = torch.zeros(cout, device=DEV)
        res = {}
        wn, nt = engine.pack_tma_weights(wt)
        y_new = engine.conv3d_tma(x, wn, nt, cout, kd, shift, None, True)
        res["tma_us"] = timed(lambda: engine.conv3d_tma(x, wn, nt, cout, kd, shift, None, True))
        hi, _, ntk = engine.pack_tc_weights(wt, False)
        old = lambda: engine.conv3d_tc(x, hi, None, ntk, cout, kd, shift, None, (1, 1, 1), True)
        y_old = old()
        res["old_us"] = timed(old)
        res["max_abs_diff"] = float((y_new.reshape(-1) - y_old.reshape(-1)).abs().max())
        flops = 2.0 * 9 * kd * cin * cout * b * d * h * w
        nbytes = 4.0 * b * d * h * w * (cin + cout)
        res["tma_tflops"] = flops / res["tma_us"] / 1e6
        res["tma_gbs"] = nbytes / res["tma_us"] / 1e3
        out[name] = res
        print(name, {k: (round(v, 2) if isinstance(v, float) else v) for k, v in res.items()}, flush=True)


if __name__ == "__main__":
    main()
