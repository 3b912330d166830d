"""world_size-2 gloo test of the sharding logic: each rank evaluates its
contiguous range of global sample ids, ONE all-reduce combines the moment buffers, and the
result equals the unsharded sum.  The per-rank compute is stood in for by the oracle here
(CPU container); the GPU version of the same test is tests/test_gpu_parity.py::test_sharded."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from deepbayes_b200 import distributed as dd
    from oracle import oracle as o
    L = o.mlp([11, 8, 4])
    mean, std = o.synthetic_posterior(L, seed=1, sigma_scale=0.4)
    x = o.synthetic_x((5, 11), seed=0)
    n = 7
    eps = o.synthetic_eps(L, n, seed=2)
    assert dd.world() == (rank, world)
    off, cnt = dd.shard_range(n, rank, world)
    s1, s2 = o.predict_moments(L, mean, std, x, cnt, eps[off:off + cnt])
    t1, t2 = torch.from_numpy(s1.copy()), torch.from_numpy(s2.copy())
    dd.allreduce_sum_(t1, t2)
    f1, f2 = o.predict_moments(L, mean, std, x, n, eps)
    ok = np.allclose(t1.numpy(), f1, rtol=1e-6, atol=1e-7) and np.allclose(t2.numpy(), f2, rtol=1e-6, atol=1e-7)
    q.put((rank, bool(ok), off, cnt))
    dist.destroy_process_group()


def test_two_rank_sharded_moments():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
    assert [r[1] for r in res] == [True, True]
    assert (res[0][2], res[0][3]) == (0, 4) and (res[1][2], res[1][3]) == (4, 3)
