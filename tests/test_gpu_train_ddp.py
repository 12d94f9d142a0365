"""Two-GPU NCCL run of the training path (BASELINE cfg 5: one reference view per rank, SyncBatchNorm,
DDP gradient all-reduce over NVLink): 2 ranks x 1 item must equal 1 rank x 2 items — the device twin of
tests/test_train_ddp_gloo.py.  Needs >= 2 GPUs (skipped on a one-GPU box)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F

from mvsformer_b200 import synthetic as S
from tests.helpers import STAGE_ARGS, rel_l1

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _inputs():
    from tests.test_train_emulated import _case

    stage, ndepth = 3, 4
    feats, cams, hyp = _case(batch=2, views=3, chans=S.FEAT_CHS[stage], depth=ndepth, height=32, width=48, seed=41)
    target = torch.randint(0, ndepth, (2, 32, 48), generator=S._gen(8))
    return stage, ndepth, feats, cams, hyp, target


def _make_net(stage, ndepth):
    from mvsformer_b200.mvsformer_model import StageNet

    net = StageNet(dict(STAGE_ARGS), ndepth, stage).train()
    net.load_state_dict(S.fill_state_dict(net.state_dict(), seed=33))
    return net


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    stage, ndepth, feats, cams, hyp, target = _inputs()
    net = torch.nn.SyncBatchNorm.convert_sync_batchnorm(_make_net(stage, ndepth)).cuda(rank)
    ddp = torch.nn.parallel.DistributedDataParallel(net, device_ids=[rank])
    f = feats[rank:rank + 1].cuda(rank).requires_grad_(True)
    res = ddp(f, cams[rank:rank + 1].cuda(rank), hyp[rank:rank + 1].contiguous().cuda(rank))
    F.cross_entropy(res["prob_volume_pre"], target[rank:rank + 1].cuda(rank)).backward()
    torch.cuda.synchronize()
    grads = {k: p.grad.cpu().numpy() for k, p in net.named_parameters()}
    gathered = [None] * world
    dist.all_gather_object(gathered, (f.grad.cpu().numpy(), res["prob_volume_pre"].detach().cpu().numpy()))
    if rank == 0:
        out.put((grads, gathered))
    dist.barrier()
    dist.destroy_process_group()


def test_syncbn_ddp_two_gpus_equals_single_gpu_batch():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    try:
        grads, gathered = q.get(timeout=300)
    finally:
        for p in procs:
            p.join(60)
            if p.is_alive():
                p.terminate()
    assert all(p.exitcode == 0 for p in procs)
    stage, ndepth, feats, cams, hyp, target = _inputs()
    net = _make_net(stage, ndepth).cuda(0)
    f = feats.cuda(0).requires_grad_(True)
    out = net(f, cams.cuda(0), hyp.cuda(0))
    F.cross_entropy(out["prob_volume_pre"], target.cuda(0)).backward()
    pre2 = torch.cat([torch.from_numpy(g[1]) for g in gathered], dim=0)
    assert rel_l1(pre2, out["prob_volume_pre"].cpu()) < 1e-4
    for k, p in net.named_parameters():
        if k == "cost_reg.prob.bias":
            continue
        assert rel_l1(torch.from_numpy(grads[k]), p.grad.cpu()) < 2e-2, k
    fgrad2 = torch.cat([torch.from_numpy(g[0]) for g in gathered], dim=0)
    assert rel_l1(fgrad2 * 0.5, f.grad.cpu()) < 5e-3
