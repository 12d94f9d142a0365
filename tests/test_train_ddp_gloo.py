"""World-size-2 gloo test (CPU) of the multi-GPU TRAINING host logic (replicas, one
reference view per rank, NCCL gradient all-reduce by DDP; SyncBatchNorm as train.py:138 converts it).

Two processes each own one batch item; the model is SyncBatchNorm-converted and wrapped in
DistributedDataParallel over gloo, with the package's kernels replaced by the CPU thread-emulation of
the same source (tests/emu).  The run must reproduce single-process training on the batch of two:
global batch statistics in the forward, all-reduced statistic gradients in the backward, DDP-averaged
parameter gradients.  (On the GPU box the same code path runs over NCCL.)"""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F

from mvsformer_b200 import synthetic as S
from tests.helpers import STAGE_ARGS, rel_l1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _inputs():
    from tests.test_train_emulated import _case

    stage, ndepth = 3, 4
    feats, cams, hyp = _case(batch=2, views=3, chans=S.FEAT_CHS[stage], depth=ndepth, height=8, width=16, seed=41)
    target = torch.randint(0, ndepth, (2, 8, 16), generator=S._gen(8))
    return stage, ndepth, feats, cams, hyp, target


def _make_net(stage, ndepth):
    from mvsformer_b200.mvsformer_model import StageNet

    net = StageNet(dict(STAGE_ARGS), ndepth, stage).train()
    net.load_state_dict(S.fill_state_dict(net.state_dict(), seed=33))
    return net


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tests.emu import harness

    harness.install(setattr)
    stage, ndepth, feats, cams, hyp, target = _inputs()
    net = _make_net(stage, ndepth)
    # DDP refuses SyncBatchNorm on CPU modules at construction, so wrap first and convert afterwards: the
    # conversion keeps the same Parameter objects (DDP's hooks stay valid); buffers are not broadcast.
    ddp = torch.nn.parallel.DistributedDataParallel(net, broadcast_buffers=False)
    torch.nn.SyncBatchNorm.convert_sync_batchnorm(net)
    assert isinstance(net.cost_reg.conv1.bn, torch.nn.SyncBatchNorm) and isinstance(net.vis[0].bn, torch.nn.SyncBatchNorm)
    f = feats[rank:rank + 1].clone().requires_grad_(True)
    out_d = ddp(f, cams[rank:rank + 1], hyp[rank:rank + 1].contiguous())
    F.cross_entropy(out_d["prob_volume_pre"], target[rank:rank + 1]).backward()
    # numpy arrays travel through the queue by value (torch tensors would need the producer to stay alive)
    grads = {k: p.grad.numpy().copy() for k, p in net.named_parameters()}
    stats = {k: b.numpy().copy() for k, b in net.named_buffers() if "running" in k}
    gathered = [None] * world
    dist.all_gather_object(gathered, (f.grad.numpy(), out_d["prob_volume_pre"].detach().numpy()))
    if rank == 0:
        out.put((grads, stats, gathered))
    dist.barrier()
    dist.destroy_process_group()


def test_syncbn_ddp_two_ranks_equals_single_process_batch(monkeypatch):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    try:
        grads, stats, gathered = q.get(timeout=240)          # before join: the payload is larger than a pipe buffer
    finally:
        for p in procs:
            p.join(60)
            if p.is_alive():
                p.terminate()
    assert all(p.exitcode == 0 for p in procs)
    # single process, batch of two, plain BatchNorm
    from tests.emu import harness

    harness.install(monkeypatch.setattr)
    stage, ndepth, feats, cams, hyp, target = _inputs()
    net = _make_net(stage, ndepth)
    f = feats.clone().requires_grad_(True)
    out = net(f, cams, hyp)
    F.cross_entropy(out["prob_volume_pre"], target).backward()
    pre2 = torch.cat([torch.from_numpy(g[1]) for g in gathered], dim=0)
    assert rel_l1(pre2, out["prob_volume_pre"]) < 1e-5                       # global batch statistics in the forward
    for k, p in net.named_parameters():
        if k == "cost_reg.prob.bias":
            continue
        assert rel_l1(grads[k], p.grad) < 1e-3, k                              # DDP mean of per-rank grads
    fgrad2 = torch.cat([torch.from_numpy(g[0]) for g in gathered], dim=0)
    assert rel_l1(fgrad2 * 0.5, f.grad) < 1e-3                                 # per-rank loss is a mean over 1 item, not 2
    for k, b in net.named_buffers():
        if "running" in k:
            assert rel_l1(stats[k], b) < 1e-5, k
