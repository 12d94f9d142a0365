"""Multi-GPU training step of the hot path at BASELINE cfg 5's shape (one reference + 4 source views of 512x640 per
GPU, 4-stage cascade, SyncBatchNorm, DistributedDataParallel over NCCL) — an extra measurement, not the headline bench.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 \\
        scripts/train_ddp_bench.py [--steps 10] [--warmup 3] [--no-sync-bn]

Every rank owns its own synthetic reference view (features resident in HBM; the backbone is out of scope), the step is
forward + CE loss on the four stages + backward (our kernels; DDP all-reduces the 302 `fusions.*` gradients, 4.7 MB) +
SGD.  Timing: barrier + synchronize on both sides, CUDA events, MAX over ranks; rank 0 prints one JSON line with the
whole-job reference views / s.  Written without GPU time in round 1: first run is round 2's.
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mvsformer_b200 import _lib, synthetic as S  # noqa: E402
from mvsformer_b200.mvsformer_model import CascadeMVS  # noqa: E402

ARGS = {"base_ch": 8, "fusion_type": "cnn", "depth_type": "ce", "ndepths": list(S.NDEPTHS),
        "depth_interals_ratio": list(S.DEPTH_INTERVAL_RATIO), "inverse_depth": True}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--no-sync-bn", action="store_true")
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/nccl_debug.%h.%p.log")
        dist.init_process_group("nccl", device_id=dev)
    height, width, views = 512, 640, 5
    feats = {k: v.to(dev).requires_grad_(True) for k, v in S.make_features(1, views, height, width, seed=100 + rank).items()}
    cams = {k: v.to(dev) for k, v in S.make_cameras(1, views, height, width).items()}
    dv = S.make_depth_range(1).to(dev)
    targets = [torch.randint(0, S.NDEPTHS[s], (1,) + S.stage_hw(height, width, s), generator=S._gen(10 * rank + s)).to(dev)
               for s in range(4)]
    net = CascadeMVS(dict(ARGS)).train()
    if world > 1 and not a.no_sync_bn:
        net = torch.nn.SyncBatchNorm.convert_sync_batchnorm(net)
    net = net.to(dev)
    model = torch.nn.parallel.DistributedDataParallel(net, device_ids=[local]) if world > 1 else net
    opt = torch.optim.SGD(net.parameters(), lr=1e-3)

    def step():
        opt.zero_grad(set_to_none=True)
        out = model(feats, cams, dv)
        loss = sum(F.cross_entropy(out["stage%d" % (s + 1)]["prob_volume_pre"], targets[s]) for s in range(4))
        loss.backward()
        opt.step()
        return loss.detach()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(a.warmup):
        step()
    barrier()
    l0 = _lib.load().mvs_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        loss = step()
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        ms_step = float(ms.item()) / a.steps
        print(json.dumps({"metric": "training reference views/s (cfg 5 shape: 5 views, 512x640, 4-stage cascade; features given)",
                          "value": world * 1e3 / ms_step, "unit": "ref views/s", "n_gpus": world, "steps": a.steps,
                          "warmup": a.warmup, "ms_per_step": ms_step, "scaling": "weak", "sync_bn": world > 1 and not a.no_sync_bn,
                          "launches_per_step_rank0": (_lib.load().mvs_launch_count() - l0) / a.steps, "loss": float(loss),
                          "dtype": "f32 (CUDA-core training kernels)", "data": "synthetic"}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
