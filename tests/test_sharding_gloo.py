"""World-size-2 gloo test (CPU) of the N>1 host logic: reference views are sharded round-robin with
no data-path collective; whole-job throughput = total units / MAX-over-ranks time."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mvsformer_b200.sharding import max_over_ranks, shard_ref_views, sum_over_ranks


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard_ref_views(49, rank, world)                 # one DTU scan = 49 reference views
    elapsed = 1.0 + 0.5 * rank                               # rank 1 is the straggler
    t_max = max_over_ranks(elapsed)
    n_total = sum_over_ranks(len(mine))
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    if rank == 0:
        out.put((gathered, t_max, n_total))
    dist.barrier()
    dist.destroy_process_group()


def test_round_robin_sharding_two_ranks():
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    gathered, t_max, n_total = q.get()
    assert sorted(gathered[0] + gathered[1]) == list(range(49))         # disjoint cover
    assert abs(len(gathered[0]) - len(gathered[1])) <= 1                # balanced
    assert t_max == 1.5 and n_total == 49.0                             # throughput = 49 / 1.5


def test_shard_edges():
    assert shard_ref_views(0, 0, 4) == []
    assert shard_ref_views(3, 3, 4) == []
    assert shard_ref_views(5, 1, 2) == [1, 3]
    assert max_over_ranks(2.5) == 2.5                                    # no process group: identity
