"""Multi-GPU sharding of the hot path: reference views are independent, so inference shards them
round-robin across ranks with NO data-path collective (the reference itself is
single-GPU, test.py:212).  The only collectives are the bookkeeping ones of a benchmark/driver:
a barrier and a MAX-reduce of per-rank device times."""
import torch
import torch.distributed as dist


def shard_ref_views(num_items, rank, world_size):
    """Indices of the (scan, ref_view) metas this rank processes: i with i % world_size == rank
    (the order DistributedSampler(shuffle=False) would give, train.py:46)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank %d outside world of %d" % (rank, world_size))
    return list(range(rank, num_items, world_size))


def max_over_ranks(value, device=None):
    """MAX-reduce of a python float over the default process group (identity when not initialised)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device=None):
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
