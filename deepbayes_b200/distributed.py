"""Multi-GPU sharding of the posterior-sample loop.

Posterior samples are independent (posteriormodel.py:95-101), so rank g of G evaluates the
global sample ids [g*n/G, (g+1)*n/G) -- Philox is keyed by the GLOBAL id, so the union of the
shards is the same set of draws for every G -- and ONE all-reduce(SUM, fp32) over the [B,C]
moment buffers combines the predictive moments (NCCL over NVLink on GPUs, gloo in CPU tests).
There is no other data-path collective.
"""
from __future__ import annotations


def _dist():
    import torch.distributed as dist
    return dist


def world():
    """(rank, world_size) of the default process group, (0, 1) when not initialised."""
    try:
        dist = _dist()
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except Exception:
        pass
    return 0, 1


def shard_range(n: int, rank: int, world_size: int):
    """Contiguous, balanced split of sample ids 0..n-1: returns (offset, count)."""
    if world_size <= 1:
        return 0, n
    base, rem = divmod(n, world_size)
    cnt = base + (1 if rank < rem else 0)
    off = rank * base + min(rank, rem)
    return off, cnt


def allreduce_sum_(*tensors):
    """In-place SUM all-reduce of the moment buffers, coalesced into one collective."""
    rank, ws = world()
    ts = [t for t in tensors if t is not None]
    if ws <= 1 or not ts:
        return tensors
    import torch
    dist = _dist()
    if len(ts) == 1:
        dist.all_reduce(ts[0], op=dist.ReduceOp.SUM)
    else:
        flat = torch.cat([t.reshape(-1) for t in ts])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        o = 0
        for t in ts:
            t.copy_(flat[o:o + t.numel()].view_as(t))
            o += t.numel()
    return tensors
