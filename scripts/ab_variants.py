"""One-process A/B harness for the opt-in variants that were written without GPU time (round 1) — meant to be
the FIRST gpurun call of the next round: it runs the gated experimental parity tests, then times every variant
with CUDA events on the bench workload, and writes one JSON.

    python scripts/ab_variants.py gpurun_out/ab            # ~40 s on a B200

Variants
  cost-volume build   two sampling passes (shipped)  vs  MVS_CV_STORE (pass A stores corr, streaming aggregation)
  3D CNN              depth-fused tcgen05 convs      vs  MVS_TCZ_KZF=1/2 (kz-fused N: one MMA per slab instead of three)
  weight gradients    8x8 register tiles (default)   vs  MVS_WGRAD_TILE=4
Each inference variant reports ms per reference view (cfg 2, inputs resident) and the per-kernel-class CUDA-event
attribution of bench.py's KernelProfiler; the training variants report ms per cfg-5-shaped training step and the
per-entry-point breakdown of scripts/train_profile.py.
"""
import contextlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))
os.chdir(ROOT)
out_dir = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/ab"
os.makedirs(out_dir, exist_ok=True)

import pytest  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from mvsformer_b200 import config, engine  # noqa: E402

_CV_STORE_DEFAULT, _TCZ_KZF_DEFAULT = config.cv_store(), config.tcz_kzf()


def time_inference(net, feats, cams, dv, steps=8, warmup=3):
    tmp = list(bench.S.EVAL_TMP)
    with torch.no_grad():
        for _ in range(warmup):
            net(feats, cams, dv, tmp=tmp)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            out = net(feats, cams, dv, tmp=tmp)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        prof = bench.KernelProfiler()
        prof.install(engine)
        for _ in range(3):
            net(feats, cams, dv, tmp=tmp)
        torch.cuda.synchronize()
        prof.uninstall(engine)
        kern = {k: round(v["ms_per_step"], 4) for k, v in prof.summary(3).items()}
    return ms, kern, out["refined_depth"].clone()


def main():
    result = {}
    # gated parity tests, one pytest run per group so that a failing variant only disqualifies itself
    groups = {"cv_store": "cv_store", "kzf": "kzf or khf", "fusion": "fusion and not epipole", "heads": "heads", "epipole": "epipole",
              "train_conv": "training_convs", "diff_warp": "diff_homo"}
    ok = {}
    with open(os.path.join(out_dir, "experimental_tests.log"), "w") as f, contextlib.redirect_stdout(f), \
            contextlib.redirect_stderr(f):
        for name, expr in groups.items():
            print("==== group %s (-k %r)" % (name, expr), flush=True)
            ok[name] = int(pytest.main(["tests/test_gpu_variants.py", "-q", "-k", expr, "-p", "no:cacheprovider"])) == 0
    result["experimental_tests_passed"] = ok
    print("experimental tests", ok, flush=True)

    device = torch.device("cuda", 0)
    config.set_conv_precision(os.environ.get("MVS_CONV_PRECISION", "tf32"))
    net = bench.build_engine(device)
    feats_h, cams_h, dv_h = bench.host_inputs(bench.HEIGHT, bench.WIDTH, bench.VIEWS, 1234, pin=False)
    feats = {k: v.to(device) for k, v in feats_h.items()}
    cams = {k: v.to(device) for k, v in cams_h.items()}
    dv = dv_h.to(device)
    depths = {}
    for name, store, kzf in (("shipped", False, 0), ("cv_store", True, 0), ("tcz_kzf_1", False, 1), ("tcz_kzf_2", False, 2),
                             ("cv_store+tcz_kzf_2", True, 2)):
        if (store and not ok["cv_store"]) or (kzf and not ok["kzf"]):
            result[name] = {"skipped": "its gated parity tests failed"}
            continue
        config.set_cv_store(store)
        config.set_tcz_kzf(kzf)
        ms, kern, depths[name] = time_inference(net, feats, cams, dv)
        result[name] = {"ms_per_ref_view": ms, "maps_per_s": 1e3 / ms, "kernels_ms": kern}
        print(name, round(ms, 3), "ms", flush=True)
    config.set_cv_store(_CV_STORE_DEFAULT)
    config.set_tcz_kzf(_TCZ_KZF_DEFAULT)
    if "cv_store" in depths:
        result["cv_store_refined_depth_bit_identical"] = bool(torch.equal(depths["shipped"], depths["cv_store"]))
    if "tcz_kzf_2" in depths:
        d0, d2 = depths["shipped"], depths["tcz_kzf_2"]
        result["tcz_kzf_refined_depth_rel_l1"] = float((d2 - d0).abs().mean() / d0.abs().mean())

    # depth-map fusion at DTU size with 10 source views: HBM-bound, (1 + V) maps in, (5 V + 2) out
    if ok["fusion"]:
        from mvsformer_b200 import fusion as Fu
        case = {k: v.to(device) for k, v in bench.S.make_fusion_case(11, bench.HEIGHT, bench.WIDTH, seed=2).items()}
        for _ in range(3):
            Fu.filter_view(case["ref_depth"], case["src_depths"], case["ref_cam"], case["src_cams"], 1.0, 0.01, 3)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            Fu.filter_view(case["ref_depth"], case["src_depths"], case["ref_cam"], case["src_cams"], 1.0, 0.01, 3)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        v, hw = 10, bench.HEIGHT * bench.WIDTH
        alg = 4 * hw * ((1 + v) + (4 * v) + (v + 2) + (1 + 3 * v + v) + 1 + 3)      # reproject in/out, filter in/out, points
        result["fusion_filter_view"] = {"ms_per_ref_view": ms, "alg_bytes": alg, "gb_per_s": alg / ms / 1e6}
        print("fusion", round(ms, 3), "ms", flush=True)

    import train_profile  # noqa: E402
    for tile in ("8", "4"):
        os.environ["MVS_WGRAD_TILE"] = tile
        path = os.path.join(out_dir, "train_profile_tile%s.json" % tile)
        with open(path, "w") as f, contextlib.redirect_stdout(f):
            train_profile.main()
        d = json.load(open(path))
        result["train_wgrad_tile" + tile] = {"step_ms": d["step_ms_unprofiled"], "kernel_ms_sum": d["kernel_ms_sum"],
                                             "entry_points_ms": {k: v["ms"] for k, v in d["entry_points"].items()}}
        print("train tile", tile, round(d["step_ms_unprofiled"], 2), "ms", flush=True)
    os.environ.pop("MVS_WGRAD_TILE", None)
    for mode in ("tf32x3", "tf32"):                       # training convs on the tensor cores (config.train_conv)
        if not ok["train_conv"]:
            result["train_conv_" + mode] = {"skipped": "its gated parity tests failed"}
            continue
        config.set_train_conv(mode)
        path = os.path.join(out_dir, "train_profile_%s.json" % mode)
        try:
            with open(path, "w") as f, contextlib.redirect_stdout(f):
                train_profile.main()
            d = json.load(open(path))
            result["train_conv_" + mode] = {"step_ms": d["step_ms_unprofiled"], "kernel_ms_sum": d["kernel_ms_sum"],
                                            "entry_points_ms": {k: v["ms"] for k, v in d["entry_points"].items()}}
        except Exception as exc:
            result["train_conv_" + mode] = {"error": "%s: %s" % (type(exc).__name__, str(exc)[:300])}
        print("train conv", mode, result["train_conv_" + mode].get("step_ms"), flush=True)
    config.set_train_conv("fp32")
    with open(os.path.join(out_dir, "ab_variants.json"), "w") as f:
        json.dump(result, f, indent=1)
    print(json.dumps(result)[:3000])


if __name__ == "__main__":
    main()
