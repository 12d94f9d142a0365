#!/usr/bin/env python
"""bench.py — depth maps/s of the plane-sweep cascade on B200 (BASELINE.json metric, cfg 2).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one reference view: the 4-stage cascade (hypothesis
schedule -> fused warp/correlation cost volume -> 3D-CNN -> softmax regression) at DTU test size
1152x1536, 5 views, 192-depth range, ndepths 32/16/8/4, synthetic features / cameras / weights.
Reference views are independent, so with N GPUs every rank runs its own reference views
(weak scaling, no data-path collective); `value` = N*K / max-over-ranks(device time).

Printed JSON line (rank 0): value (inputs resident in HBM), e2e (the public host-to-host call over a scan: pinned
HOST features in, HOST depth + confidence out, H2D + re-layout + D2H inside the timed region; consecutive reference
views share 4 of their 5 views as in a DTU scan, so each view crosses PCIe once — e2e_dense re-uploads all five views
of every reference view), rooflines (live CUDA-event timing per kernel class), cpu_baseline, gpu_eager_baseline, clocks.

`--impl reference` times the UNMODIFIED reference's hot path (its own StageNet modules and schedulers, staged
under oracle/_ref by oracle/build_ref.py; the oracle port only if the staged copy is missing) on the host cores:
every step is a bounded 256x384 sample of the workload (rate scaled by the pixel ratio, which the line's own
`config` says), followed by ONE untimed-by-the-driver full-size cfg-2 pass that checks the extrapolation.
`--impl eager` (and the `gpu_eager_baseline` key of the engine line) runs the same reference modules on cuda:0:
PyTorch eager + cuDNN on the same B200, the baseline the hand-written kernels have to beat.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from mvsformer_b200 import synthetic as S  # noqa: E402

METRIC = "depth maps/sec @ DTU 1152x1536, 5 views, 192 depths; cost-vol Gvox/s"
UNIT = "depth maps/s"
HEIGHT, WIDTH, VIEWS = 1152, 1536, 5
SAMPLE_H, SAMPLE_W = 256, 384          # bounded CPU sample (1/18 of the pixels of the full workload)
CASCADE_ARGS = {"base_ch": 8, "fusion_type": "cnn", "depth_type": "ce", "ndepths": list(S.NDEPTHS),
                "depth_interals_ratio": list(S.DEPTH_INTERVAL_RATIO), "inverse_depth": True}


def conv_mode():
    from mvsformer_b200 import config
    return {"tf32x3": "3xTF32 tcgen05 (fp32-grade)", "tf32": "TF32 tcgen05", "fp32": "FP32 CUDA cores"}[config.conv_precision()]


def workload_config(n_gpus):
    from mvsformer_b200 import config
    return {"cost_volume_build": "channels-last kernels, one sampling pass at stages 1-3" if config.cv_layout() == "cl"
            else "generic NCHW kernels, two sampling passes", "workload": "DTU test (cfg 2): %dx%d, %d views, 192-depth range, 4-stage cascade ndepths 32/16/8/4, "
                        "feat ch 64/32/16/8, G=8, B=1 ref view per GPU per step" % (HEIGHT, WIDTH, VIEWS),
            "precision": "fp32 features/volume/activations; conv math %s" % conv_mode(), "parallelism": "ref views sharded, %d rank(s), no collective" % n_gpus,
            "l2_policy": "inputs 530 MB/step > 126 MB L2; every intermediate volume is rewritten each step"}


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# per-kernel-class timing (CUDA events on torch's current stream = the launching stream)
# ------------------------------------------------------------------------------------------------
class KernelProfiler:
    """Wraps the engine entry points with CUDA events and algorithmic byte / flop counters."""

    def __init__(self):
        self.records = {}
        self._saved = {}

    def _wrap(self, engine, name, cls, cost):
        fn = getattr(engine, name)
        self._saved[name] = fn

        def wrapped(*a, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn(*a, **k)
            e1.record()
            nbytes, flops = cost(out, *a, **k)
            name_cls = cls(*a, **k) if callable(cls) else cls
            rec = self.records.setdefault(name_cls, {"events": [], "bytes": 0, "flops": 0, "launches": 0})
            rec["events"].append((e0, e1))
            rec["bytes"] += nbytes
            rec["flops"] += flops
            rec["launches"] += 1
            return out

        setattr(engine, name, wrapped)

    def install(self, engine):
        def numel_bytes(*ts):
            return sum(4 * t.numel() for t in ts if t is not None and hasattr(t, "numel"))

        def cv_cost(out, features, relproj, depth_values, *rest, **k):
            b, v, c, h, w = features.shape
            d = depth_values.shape[1]
            return 4 * (b * v * c * h * w + b * d * h * w), 0      # features + hypotheses (outputs added below)

        def ent_cost(out, features, relproj, depth_values, groups, want_sim):
            base, _ = cv_cost(out, features, relproj, depth_values)
            return base + numel_bytes(out[0], out[1]), 0

        def agg_cost(out, features, relproj, depth_values, vis_weight, groups, round_tf32=False):
            base, _ = cv_cost(out, features, relproj, depth_values)
            return base + numel_bytes(out, vis_weight), 0

        def conv_cost(out, x, w_packed, shift, skip, stride, relu=True):
            taps_cin_cout = w_packed.numel()
            return numel_bytes(x, out, skip, w_packed), 2 * taps_cin_cout * (out.numel() // out.shape[-1])

        def deconv_cost(out, x, w_packed, shift, skip, sd, relu=True):
            return numel_bytes(x, out, skip, w_packed), 2 * w_packed.numel() * (x.numel() // x.shape[-1])

        def vis_cost(out, ent, params):
            return numel_bytes(ent, out), 2 * 3608 * ent.numel()

        def io_cost(out, *a, **k):
            outs = out if isinstance(out, (tuple, list)) else (out,)
            return numel_bytes(*outs) + numel_bytes(*[t for t in a if torch.is_tensor(t)]), 0

        def cl_ent_cost(out, feat_cl, relproj, depth_values, groups, want_sim):
            if out is None:
                return 0, 0
            return numel_bytes(feat_cl, depth_values, *out), 0

        def cl_agg_cost(out, feat_cl, relproj, depth_values, vis_weight, groups, round_tf32=False):
            return numel_bytes(feat_cl, depth_values, vis_weight, out), 0

        def to_cl_cost(out, feature_list, outs=None):
            return 2 * numel_bytes(*feature_list), 0

        # channels-last cost-volume kernels (csrc/cost_volume_cl.cu)
        self._wrap(engine, "features_to_cl", "cv_layout(nchw->channels-last)", to_cl_cost)
        self._wrap(engine, "cost_volume_cl_entropy", lambda f, r, d, g, want_sim: "cv_cl_passA+store" if f.shape[-1] >= 16 else "cv_cl_passA(stage4)",
                   cl_ent_cost)
        self._wrap(engine, "cost_volume_cl_aggregate", "cv_cl_passB(stage4)", cl_agg_cost)
        self._wrap(engine, "cost_volume_entropy", "cv_entropy(passA)", ent_cost)
        self._wrap(engine, "cost_volume_aggregate", "cv_aggregate(passB)", agg_cost)
        self._wrap(engine, "corr_aggregate", "cv_corr_aggregate(stream)", io_cost)
        def vis_fused_cost(out, ent, params, w2, w3):
            return numel_bytes(ent, out), 2 * 3608 * ent.numel()

        self._wrap(engine, "vis_fused", "vis_net(fused)", vis_fused_cost)
        self._wrap(engine, "vis_weight", "vis_net", vis_cost)
        def conv_tc_cost(out, x, w_hi, w_lo, n_tile, cout, kd, shift, skip, stride, relu=True):
            return numel_bytes(x, out, skip), 2 * kd * 9 * x.shape[-1] * cout * (out.numel() // out.shape[-1])

        def deconv_tc_cost(out, x, w_hi, w_lo, n_tile, cout, kd, shift, skip, sd, relu=True):
            return numel_bytes(x, out, skip), 2 * kd * 9 * x.shape[-1] * cout * (x.numel() // x.shape[-1])

        def conv_tma_cost(out, x, w_tma, n_tile, cout, kd, shift, skip, relu=True, mode=0):
            vox = (x.numel() // x.shape[-1]) if mode == 2 else (out.numel() // out.shape[-1])
            return numel_bytes(x, out, skip), 2 * kd * 9 * x.shape[-1] * cout * vox

        def conv_tma_prob_cost(out, x, w_tma, kd, shift, skip, prob_w_host, prob_bias, relu=True):
            vox = x.numel() // x.shape[-1]                                                # transposed 16 -> 8 + the 1x1x1 prob conv
            return numel_bytes(x, out, skip), 2 * kd * 9 * 16 * 8 * vox + 2 * 8 * out.numel()

        self._wrap(engine, "conv3d_tma", "conv3d_tma", conv_tma_cost)                    # persistent TMA-fed kernels
        self._wrap(engine, "conv3d_tma_prob", "conv3d_tma", conv_tma_prob_cost)
        self._wrap(engine, "conv3d_cl", "conv3d", conv_cost)
        self._wrap(engine, "deconv3d_cl", "deconv3d", deconv_cost)
        self._wrap(engine, "conv3d_tc", "conv3d_tc", conv_tc_cost)
        self._wrap(engine, "deconv3d_tc", "deconv3d_tc", deconv_tc_cost)
        for name in ("prob_conv_cl", "regression_head", "argmax_gather", "init_range", "schedule_inverse_range",
                     "confidence_accumulate", "confidence_upsample_accumulate", "relative_projections"):
            self._wrap(engine, name, "head+schedule", io_cost)

    def uninstall(self, engine):
        for name, fn in self._saved.items():
            setattr(engine, name, fn)

    def summary(self, steps):
        out = {}
        for cls, rec in self.records.items():
            ms = sum(e0.elapsed_time(e1) for e0, e1 in rec["events"])
            out[cls] = {"ms_per_step": ms / steps, "launches_per_step": rec["launches"] / steps,
                        "alg_bytes_per_step": rec["bytes"] / steps, "alg_flops_per_step": rec["flops"] / steps,
                        "avg_launch_ms": ms / max(rec["launches"], 1)}
        return out


# ------------------------------------------------------------------------------------------------
# workload construction
# ------------------------------------------------------------------------------------------------
def build_engine(device):
    from mvsformer_b200.mvsformer_model import CascadeMVS
    net = CascadeMVS(dict(CASCADE_ARGS)).eval()
    full = {}
    for s in range(4):
        sd = S.fill_state_dict(net.fusions[s].state_dict(), seed=s)
        full.update({"fusions.%d.%s" % (s, k): v for k, v in sd.items()})
    net.load_state_dict(full, strict=True)
    return net.to(device)


def host_inputs(height, width, views, seed, pin):
    feats = S.make_features(1, views, height, width, seed=seed)
    cams = S.make_cameras(1, views, height, width)
    dv = S.make_depth_range(1)
    if pin:
        feats = {k: v.pin_memory() for k, v in feats.items()}
        cams = {k: v.pin_memory() for k, v in cams.items()}
        dv = dv.pin_memory()
    return feats, cams, dv


def best_thread_count():
    """The oracle port is many small torch-CPU ops; on a many-core host the OpenMP fork/join of
    all cores can cost more than it buys.  Calibrate on a 128x192 cascade and keep the fastest
    thread count (this is "all the host threads it can use")."""
    from mvsformer_b200.mvsformer_model import CascadeMVS
    from oracle import mvs_oracle as O
    ncpu = os.cpu_count() or 1
    cands = sorted({c for c in (8, 16, 32, 64, ncpu) if c <= ncpu})
    feats, cams, dv = host_inputs(128, 192, VIEWS, 7, pin=False)
    net = CascadeMVS(dict(CASCADE_ARGS)).eval()
    sds = [S.fill_state_dict(net.fusions[s].state_dict(), seed=s) for s in range(4)]
    best, best_t = cands[0], float("inf")
    with torch.no_grad():
        for c in cands:
            torch.set_num_threads(c)
            O.cascade_forward(feats, cams, dv, sds)
            t0 = time.perf_counter()
            O.cascade_forward(feats, cams, dv, sds)
            dt = time.perf_counter() - t0
            if dt < best_t:
                best, best_t = c, dt
    return best


def bench_state_dicts():
    from mvsformer_b200.mvsformer_model import CascadeMVS
    net = CascadeMVS(dict(CASCADE_ARGS)).eval()
    return [S.fill_state_dict(net.fusions[s].state_dict(), seed=s) for s in range(4)]


def cpu_kind():
    from oracle import ref_cascade
    return "reference" if ref_cascade.available() else "port"


def cpu_cascade_rate(steps, warmup, threads, height=SAMPLE_H, width=SAMPLE_W):
    """The reference's own hot path on the host (the unmodified modules staged under oracle/_ref; the oracle port
    when they are missing) on a height x width workload; returns (maps/s scaled to the full workload by the
    pixel ratio, seconds per step)."""
    from oracle import mvs_oracle as O
    from oracle import ref_cascade
    torch.set_num_threads(threads)
    feats, cams, dv = host_inputs(height, width, VIEWS, 1234, pin=False)
    sds = bench_state_dicts()
    if ref_cascade.available():
        nets = ref_cascade.build_stage_nets(sds)
        dt = ref_cascade.time_cpu(nets, feats, cams, dv, steps, warmup)
    else:
        with torch.no_grad():
            for _ in range(warmup):
                O.cascade_forward(feats, cams, dv, sds)
            t0 = time.perf_counter()
            for _ in range(steps):
                O.cascade_forward(feats, cams, dv, sds)
            dt = (time.perf_counter() - t0) / steps
    frac = (height * width) / float(HEIGHT * WIDTH)
    return frac / dt, dt


def gpu_eager_baseline(device, steps=10, warmup=3):
    """The reference's own modules on the GPU: PyTorch eager + cuDNN (TF32 convs as torch defaults), cfg 2 full size,
    inputs resident, CUDA events.  None when oracle/_ref is not staged."""
    from oracle import ref_cascade
    if not ref_cascade.available():
        return None
    feats, cams, dv = host_inputs(HEIGHT, WIDTH, VIEWS, 1234, pin=False)
    feats = {k: v.to(device) for k, v in feats.items()}
    cams = {k: v.to(device) for k, v in cams.items()}
    nets = ref_cascade.build_stage_nets(bench_state_dicts(), device)
    t = ref_cascade.time_cuda(nets, feats, cams, dv.to(device), steps, warmup)
    out = {"value": 1e3 / t["ms_per_step"], "unit": UNIT, "ms_per_step": t["ms_per_step"],
           "parts_ms": {"cnn(cost_reg x4)": t["cnn_ms"], "cost_volume_build+head": t["cost_volume_and_head_ms"]},
           "steps": steps, "warmup": warmup,
           "kind": "unmodified reference StageNet/scheduler modules (oracle/_ref) on the same GPU: PyTorch %s eager + cuDNN, "
                   "cudnn.allow_tf32=%s, matmul.allow_tf32=%s, inputs resident" % (torch.__version__, torch.backends.cudnn.allow_tf32,
                                                                                    torch.backends.cuda.matmul.allow_tf32)}
    del nets, feats
    torch.cuda.empty_cache()
    return out


def full_size_parity(net, feats_d, cams_d, dv_d, feats_h, cams_h, dv_h, tmp):
    """Refined depth of the precision mode being timed vs the fp32 CPU oracle on the SAME full-size cfg-2 inputs."""
    from oracle import mvs_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    with torch.no_grad():
        got = net(feats_d, cams_d, dv_d, tmp=tmp)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        want = O.cascade_forward({k: v.clone() for k, v in feats_h.items()}, cams_h, dv_h, bench_state_dicts())
        dt = time.perf_counter() - t0
    d, c = got["refined_depth"].cpu(), got["photometric_confidence"].cpu()
    rel = float((d - want["refined_depth"]).abs().mean() / want["refined_depth"].abs().mean())
    relc = float((c - want["photometric_confidence"]).abs().mean() / want["photometric_confidence"].abs().mean())
    return {"refined_depth_rel_l1": rel, "confidence_rel_l1": relc, "tolerance": 1e-3, "ok": bool(rel < 1e-3),
            "against": "oracle/mvs_oracle.py (fp32 torch-CPU restatement, pinned to the reference's goldens) on the full cfg-2 inputs of this run",
            "oracle_seconds": dt}


def sample_desc(height=SAMPLE_H, width=SAMPLE_W):
    if (height, width) == (HEIGHT, WIDTH):
        return "the full cfg-2 workload (%dx%d, %d views, 4-stage cascade), no extrapolation" % (HEIGHT, WIDTH, VIEWS)
    return ("4-stage cascade, %d views, 192-depth range, image %dx%d = 1/%d of the %dx%d pixels; rate scaled by the "
            "pixel ratio" % (VIEWS, height, width, round(HEIGHT * WIDTH / (height * width)), HEIGHT, WIDTH))


def cpu_threads():
    """All host cores for the unmodified reference (what a user of the reference gets); the calibrated best count
    for the oracle port (many tiny ops: fork/join of every core can cost more than it buys)."""
    return (os.cpu_count() or 1) if cpu_kind() == "reference" else best_thread_count()


def bind_to_gpu_numa_node(index):
    """Best effort: run this rank (and first-touch its pinned buffers) on the CPUs of the GPU's NUMA node, so that N ranks do
    not all stream their uploads out of node 0 (round-1 e2e scaling: 0.41 at N = 8).  Returns the node, or a string saying
    why nothing was bound (the gpurun boxes expose no NUMA topology to the container: sysfs reports node -1)."""
    try:
        bus = subprocess.run(["nvidia-smi", "-i", str(index), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=10).stdout.strip().lower()
        if bus.startswith("00000000:"):
            bus = bus[4:]
        with open("/sys/bus/pci/devices/%s/numa_node" % bus) as f:
            node = int(f.read().strip())
        if node < 0:
            return "not bound: /sys/bus/pci/devices/%s/numa_node = %d (no NUMA topology exposed)" % (bus, node)
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            cpus = set()
            for part in f.read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        allowed = cpus & set(os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
        return node
    except Exception as exc:
        return "not bound: %s" % type(exc).__name__


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1400.0, "source": "fallback"}


# ------------------------------------------------------------------------------------------------
def reference_config(n_gpus, height, width):
    return {"workload": "DTU test (cfg 2): %dx%d, %d views, 192-depth range, 4-stage cascade ndepths 32/16/8/4, feat ch 64/32/16/8, G=8; "
                        "each timed step = %s" % (HEIGHT, WIDTH, VIEWS, sample_desc(height, width)),
            "precision": "fp32 torch CPU kernels (the reference's own arithmetic)",
            "parallelism": "rank 0 only, host cores; %d GPU(s) idle" % n_gpus}


def run_reference(args, rank, world):
    if rank != 0:
        return
    kind, threads = cpu_kind(), cpu_threads()
    value, dt = cpu_cascade_rate(args.steps, max(args.warmup, 1), threads)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": reference_config(args.gpus, SAMPLE_H, SAMPLE_W),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "host_cpus": os.cpu_count(), "kind": kind, "sample": sample_desc()},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    if not args.no_full_size:
        # one pass over the FULL workload (outside the K timed steps): the extrapolation above against a measurement
        vf, dtf = cpu_cascade_rate(1, 0, threads, HEIGHT, WIDTH)
        line["full_size_check"] = {"value": vf, "unit": UNIT, "seconds_per_step": dtf, "steps": 1, "sample": sample_desc(HEIGHT, WIDTH),
                                   "extrapolated_over_measured": value / vf}
    print(json.dumps(line))


def run_eager(args, rank, world, local_rank):
    """`--impl eager`: the reference's modules on cuda (one line per invocation, rank 0 only)."""
    if rank != 0:
        return
    torch.cuda.set_device(local_rank)
    g = gpu_eager_baseline(torch.device("cuda", local_rank), args.steps, max(args.warmup, 3))
    if g is None:
        print(json.dumps({"impl": "eager", "unavailable": "oracle/_ref is not staged (run python -m oracle.build_ref where /root/reference exists)"}))
        return
    print(json.dumps({"impl": "eager", "metric": METRIC, "value": g["value"], "unit": UNIT, "n_gpus": 1, "steps": args.steps,
                      "warmup": max(args.warmup, 3), "ms_per_step": g["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                      "vs_baseline": None, "dtype": "f32 (cuDNN TF32 conv math)", "data": "synthetic",
                      "config": {"workload": workload_config(1)["workload"], "precision": g["kind"]}, "parts_ms": g["parts_ms"]}))


def measure_train_step(device, steps=5, warmup=3):
    """Extra key (NOT the headline): one training step at BASELINE cfg 5's per-GPU shape — 1 reference + 4
    source views, 512x640, 4-stage cascade in .train() (batch-statistics BN), cross-entropy on every stage's
    prob_volume_pre, backward through the hand-written kernels (mvsformer_b200/autograd.py), SGD update.
    Synthetic features / targets resident in HBM; CUDA events; FP32 CUDA-core first version."""
    import torch.nn.functional as F
    from mvsformer_b200.mvsformer_model import CascadeMVS

    height, width, views = 512, 640, 5
    feats = {k: v.to(device).requires_grad_(True) for k, v in S.make_features(1, views, height, width, seed=5).items()}
    cams = {k: v.to(device) for k, v in S.make_cameras(1, views, height, width).items()}
    dv = S.make_depth_range(1).to(device)
    net = CascadeMVS(dict(CASCADE_ARGS)).train().to(device)
    targets = [torch.randint(0, S.NDEPTHS[s], (1,) + S.stage_hw(height, width, s), generator=S._gen(s)).to(device)
               for s in range(4)]
    opt = torch.optim.SGD(net.parameters(), lr=1e-3)

    def step():
        opt.zero_grad(set_to_none=True)
        for f in feats.values():
            f.grad = None
        out = net(feats, cams, dv)
        loss = sum(F.cross_entropy(out["stage%d" % (s + 1)]["prob_volume_pre"], targets[s]) for s in range(4))
        loss.backward()
        opt.step()
        return loss.detach()

    from mvsformer_b200 import _lib
    first = None
    for i in range(warmup):
        loss = step()
        first = float(loss) if first is None else first
    torch.cuda.synchronize()
    l0 = _lib.load().mvs_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"ms_per_step": ms, "ref_views_per_s": 1e3 / ms, "steps": steps, "warmup": warmup,
            "launches_per_step": (_lib.load().mvs_launch_count() - l0) / steps,
            "loss_first": first, "loss_last": float(loss),
            "config": "cfg 5 per-GPU shape: B=1, 5 views, 512x640, 4-stage cascade, train mode, CE loss on 4 stages, "
                      "SGD; fp32 CUDA-core kernels; features given (backbone out of scope)"}


def profile_traffic():
    """DRAM bytes from the committed ncu --set full capture of this round (profiles/r02_traffic.json, written by
    scripts/summarise_ncu.py); {} when not captured."""
    path = os.path.join(ROOT, "profiles", "r02_traffic.json")
    if not os.path.exists(path):
        return {}
    with open(path) as f:
        return json.load(f)


def run_engine(args, rank, world, local_rank):
    import torch.distributed as dist
    from mvsformer_b200 import _lib, config, engine
    from mvsformer_b200.sharding import shard_ref_views

    # bench default: TF32 tensor-core convolutions (cuDNN's default conv math for the reference on a
    # GPU); tests/test_gpu_tensorcore.py::test_cascade_tf32_meets_north_star_tolerance gates it.
    config.set_conv_precision(os.environ.get("MVS_CONV_PRECISION", "tf32"))

    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    numa = bind_to_gpu_numa_node(local_rank)        # pinned host buffers are allocated (first touched) after this
    if world > 1:
        # keep stdout to the single JSON line: NCCL's optional version/debug banner goes to a file
        os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/nccl_debug.%h.%p.log")
        dist.init_process_group("nccl", device_id=device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    net = build_engine(device)
    # each rank owns the reference views i % world == rank of a virtual list; no data-path collective
    my_views = shard_ref_views(world * (args.steps + args.warmup), rank, world)
    feats_h, cams_h, dv_h = host_inputs(HEIGHT, WIDTH, VIEWS, 1234 + my_views[0], pin=True)
    feats_d = {k: v.to(device) for k, v in feats_h.items()}
    cams_d = {k: v.to(device) for k, v in cams_h.items()}
    dv_d = dv_h.to(device)
    tmp = list(S.EVAL_TMP)
    lib = _lib.load()

    def step_single():
        return net(feats_d, cams_d, dv_d, tmp=tmp)

    from mvsformer_b200.pipeline import CascadeLanes, PackedSample, StreamedCascade
    # --lanes N: N reference views in flight per GPU, round-robin over N compute streams (the narrow kernels of one view run
    # under the wide kernels of another); every view is still one full cascade, the step count is the number of views
    lanes = CascadeLanes(net, device, args.lanes) if args.lanes > 1 else None

    def step_resident():
        return lanes.submit(feats_d, cams_d, dv_d, tmp=tmp)[0] if lanes is not None else step_single()

    streamer = StreamedCascade(net, device, tmp, lanes=args.lanes)
    packed = PackedSample(feats_h, cams_h, dv_h)          # one pinned buffer -> one DMA per reference view

    def run_e2e(steps):
        """The public host-to-host call: pinned HOST features in, HOST depth + confidence out;
        uploads of view i+1 and downloads of view i-1 overlap the compute of view i."""
        checksum = 0.0
        for depth_h, conf_h in streamer.run(packed for _ in range(steps)):
            checksum += float(depth_h[0, 0, 0])          # the caller touches every result
        return checksum

    from mvsformer_b200.pipeline import ScanSample
    view_host = [{k: v[0, j] for k, v in feats_h.items()} for j in range(VIEWS)]     # pinned per-view slices
    scan_pos = [0]

    def run_e2e_scan(steps):
        """Scan mode: consecutive reference views share 4 of their 5 views, as in
        a DTU scan; the per-view feature cache uploads only the view that is new to the GPU."""
        def samples():
            for _ in range(steps):
                n = scan_pos[0]
                scan_pos[0] += 1
                yield ScanSample([n + j for j in range(VIEWS)], lambda vid: view_host[vid % VIEWS], cams_h, dv_h)
        checksum = 0.0
        for depth_h, conf_h in streamer.run_scan(samples(), capacity=16, keep_cache=True):     # warm-up and timed calls continue ONE scan
            checksum += float(depth_h[0, 0, 0])
        return checksum

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = lib.mvs_launch_count()
        e0.record()
        for _ in range(steps):
            fn()
        if lanes is not None:
            lanes.join()                       # the timing stream waits for every lane
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), lib.mvs_launch_count() - l0

    with torch.no_grad():
        for _ in range(args.warmup):
            step_single()
        ms_single, _ = timed(step_single, args.steps)           # one view at a time: the latency of a reference view
        for _ in range(max(args.warmup, 2 * args.lanes)):       # every lane's allocator pool reaches its steady state
            step_resident()
        if lanes is not None:
            lanes.join()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        ms_total, launches = timed(step_resident, args.steps)
        clocks = sampler.stop() if rank == 0 else None
        run_e2e(max(3, args.warmup))
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run_e2e(args.steps)
        e1.record()
        barrier()
        ms_t = torch.tensor([e0.elapsed_time(e1)], device=device)
        if world > 1:
            dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
        ms_e2e = float(ms_t.item())
        h2d, d2h = streamer.h2d_bytes, streamer.d2h_bytes

        run_e2e_scan(max(3, args.warmup))
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        streamer.host_wait_s, streamer.host_steps = 0.0, 0
        e0.record()
        run_e2e_scan(args.steps)
        e1.record()
        barrier()
        host_wait_ms = 1e3 * streamer.host_wait_s / max(streamer.host_steps, 1)
        ms_t = torch.tensor([e0.elapsed_time(e1)], device=device)
        if world > 1:
            dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
        ms_scan = float(ms_t.item())
        h2d_scan = streamer.h2d_bytes

        # per-kernel-class attribution (separate pass, same inputs)
        prof = KernelProfiler()
        prof.install(engine)
        barrier()
        psteps = min(args.steps, 5)
        for _ in range(psteps):
            step_single()                      # one view at a time: per-kernel times are not inflated by a concurrent view
        barrier()
        prof.uninstall(engine)
        kernels = prof.summary(psteps)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    ms_step = ms_total / args.steps
    value = world * args.steps / (ms_total * 1e-3)
    e2e_value = world * args.steps / (ms_e2e * 1e-3)
    peaks = measured_peaks()

    # ---- rooflines ---------------------------------------------------------------------------------
    # K1 = the cost-volume build's own kernels (warp + correlation sampling passes and the streaming aggregation), all
    # four stages.  Numerator: §8d's algorithmic bytes ONCE per reference view (features, hypotheses, volume; 1 009 MB at
    # cfg 2) — not once per pass.  `cost_volume` below adds the visibility net's time, as §8d's Gvox/s definition does.
    step_ms_sum = sum(k["ms_per_step"] for k in kernels.values())
    cv_bytes = S.cost_volume_algorithmic_bytes(VIEWS, HEIGHT, WIDTH)
    k1 = [k for n, k in kernels.items() if n.startswith("cv_")]
    k1_ms = sum(k["ms_per_step"] for k in k1)
    k1_launches = sum(k["launches_per_step"] for k in k1)
    traffic = profile_traffic()
    achieved = cv_bytes / (k1_ms * 1e-3) / 1e9
    roof = {"kernel": "cost-volume build K1: " + " + ".join(sorted(n for n in kernels if n.startswith("cv_"))), "bound": "hbm",
            "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
            "traffic": traffic.get("k1_dram_bytes_per_launch"), "alg_bytes_per_launch": cv_bytes / max(k1_launches, 1),
            "alg_bytes_per_step": cv_bytes, "ms_per_step": k1_ms, "launches_per_step": k1_launches,
            "avg_launch_ms": k1_ms / max(k1_launches, 1), "share_of_step": k1_ms / step_ms_sum,
            "note": "peak = %s copy bandwidth; numerator = SURVEY §8d bytes counted once per reference view; traffic = ncu dram "
                    "bytes per launch (profiles/r02_traffic.json)" % peaks["source"]}
    conv = [k for n, k in kernels.items() if n.startswith("conv3d") or n.startswith("deconv3d")]
    conv_ms = sum(k["ms_per_step"] for k in conv)
    conv_flops = sum(k["alg_flops_per_step"] for k in conv)
    conv_bytes = sum(k["alg_bytes_per_step"] for k in conv)
    tf32_peak = peaks["bf16_tflops"] / 2.0
    roof_conv = {"kernel": "3D-CNN convolutions (tcgen05 implicit GEMM)", "bound": "tensor", "achieved": conv_flops / (conv_ms * 1e-3) / 1e12,
                 "peak": tf32_peak, "unit": "TFLOP/s", "frac": conv_flops / (conv_ms * 1e-3) / 1e12 / tf32_peak,
                 "hbm": {"alg_bytes_per_step": conv_bytes, "achieved": conv_bytes / (conv_ms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"],
                         "frac": conv_bytes / (conv_ms * 1e-3) / 1e9 / peaks["hbm_gbs"], "unit": "GB/s"},
                 "traffic": traffic.get("conv_dram_bytes_per_step"), "ms_per_step": conv_ms, "share_of_step": conv_ms / step_ms_sum,
                 "note": "dense TF32 peak taken as half of the %s sustained cuBLAS bf16 figure; layer-wise bytes = in + out + skip per layer" % peaks["source"]}

    # cost-volume build incl. the visibility net against the HBM roofline — §8d's Gvox/s
    cv_ms = sum(kernels[k]["ms_per_step"] for k in kernels if k.startswith("cv_") or k.startswith("vis_net"))
    cost_volume = {"ms_per_step": cv_ms, "gvox_per_s": S.voxels_per_ref_view(HEIGHT, WIDTH) / (cv_ms * 1e-3) / 1e9,
                   "alg_bytes": cv_bytes, "achieved_gbs": cv_bytes / (cv_ms * 1e-3) / 1e9,
                   "frac_of_hbm": cv_bytes / (cv_ms * 1e-3) / 1e9 / peaks["hbm_gbs"]}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if config.conv_precision() == "fp32" else "f32 (conv MMA operands %s, fp32 accumulate)" % config.conv_precision(),
            "data": "synthetic", "config": dict(workload_config(world), views_in_flight_per_gpu=args.lanes),
            "e2e": {"value": world * args.steps / (ms_scan * 1e-3), "unit": UNIT, "ms_per_step": ms_scan / args.steps,
                    "h2d_bytes_per_step": h2d_scan, "d2h_bytes_per_step": d2h,
                    "h2d_gbs_aggregate": world * h2d_scan / (ms_scan / args.steps * 1e-3) / 1e9,
                    "host_blocked_on_results_ms_per_step": host_wait_ms,     # ~0 would mean the host paces the loop, not the GPU
                    "api": "mvsformer_b200.pipeline.StreamedCascade.run_scan: pinned host features of a scan in, host depth + "
                           "confidence out; consecutive reference views share 4 of their 5 views (DTU pairing), every view "
                           "crosses PCIe once, is re-laid out channels-last on the device once, and is addressed in place by "
                           "its pool slot (no gather); uploads / downloads overlap the previous / next view's cascade"},
            "e2e_dense": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                          "ms_per_step": ms_e2e / args.steps,
                          "h2d_gbs_aggregate": world * h2d / (ms_e2e / args.steps * 1e-3) / 1e9,
                          "api": "mvsformer_b200.pipeline.StreamedCascade.run: every reference view uploads all five views "
                                 "(531 MB, PCIe-bound) — what a caller without scan structure gets"},
            "single_view": {"ms_per_step": ms_single / args.steps, "value": world * args.steps / (ms_single * 1e-3), "unit": UNIT,
                            "note": "one reference view in flight (--lanes 1): the latency of a view; `value` keeps "
                                    "views_in_flight_per_gpu views in flight on as many CUDA streams"},
            "numa_node": numa, "gpu_launches": int(launches) * world, "clocks": clocks, "roofline": roof, "roofline_conv": roof_conv,
            "cost_volume": cost_volume, "kernels": kernels}
    if world == 1 and not args.no_parity:
        line["parity"] = full_size_parity(net, feats_d, cams_d, dv_d, feats_h, cams_h, dv_h, tmp)
    if world == 1 and not args.no_cpu_baseline:
        threads = cpu_threads()
        v, dt = cpu_cascade_rate(3, 1, threads)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "host_cpus": os.cpu_count(), "kind": cpu_kind(),
                                "sample": sample_desc(), "seconds_per_sample_step": dt}
    if world == 1 and not args.no_eager:
        try:
            line["gpu_eager_baseline"] = gpu_eager_baseline(device)
        except Exception as exc:
            line["gpu_eager_baseline"] = {"error": "%s: %s" % (type(exc).__name__, str(exc)[:300])}
    if world == 1 and not args.no_train_step:
        try:                                   # last on purpose: nothing above depends on it
            line["train_step_cfg5"] = measure_train_step(device)
        except Exception as exc:               # report, never lose the headline line
            line["train_step_cfg5"] = {"error": "%s: %s" % (type(exc).__name__, str(exc)[:300])}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference", "eager"])
    ap.add_argument("--lanes", type=int, default=3, help="reference views in flight per GPU (compute streams); B200 A/B: "
                    "1 -> 220.8, 2 -> 232, 3 -> 247.0 maps/s (profiles/r02i_lanes_ab.json); final build 1 -> 238.6, "
                    "3 -> 271-278 (profiles/r02k_ab.json, r02o_bench_20steps.json)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-eager", action="store_true", help="skip the gpu_eager_baseline leg (reference modules on cuda:0)")
    ap.add_argument("--no-parity", action="store_true", help="skip the full-size parity check against the CPU oracle")
    ap.add_argument("--no-full-size", action="store_true", help="reference arm: skip the one full-size pass")
    ap.add_argument("--no-train-step", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "engine" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.impl == "eager":
        run_eager(args, rank, world, local_rank)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU path (use --impl reference for the CPU arm)")
    if world != args.gpus:
        if args.gpus > 1 and world == 1:
            # convenience: re-launch under torchrun
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
                   "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.abspath(__file__)] + sys.argv[1:]
            raise SystemExit(subprocess.call(cmd))
    run_engine(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
