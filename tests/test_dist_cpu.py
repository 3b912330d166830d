"""world_size-2 gloo test of the sharding logic: each rank evaluates its
contiguous range of global sample ids, ONE all-reduce combines the moment buffers, and the
result equals the unsharded sum.  The per-rank compute is stood in for by the oracle here
(CPU container); on GPUs the same split is checked by tests/test_gpu_baseline_shapes.py::test_rank_shards_equal_one_rank
(one GPU) and by bench.py under torchrun ("shard_check": N ranks == 1 rank on the same global sample ids)."""
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
    # the drop-in chernoff_bound_verification with model.shard: the host logic (global sample ids split over the ranks, one
    # all-reduce of the [B,C] buffer) around a stand-in engine that evaluates each rank's ids with the oracle
    from deepbayes_b200.analyzers import verifiers

    class _Eng:
        K, C, device = 11, 4, torch.device("cpu")
        calls = []

        def __init__(self):
            self.torch = torch

        def ibp_predict(self, x2, e, labels, cnt_, seed, s0, noise, inflate, mode, want=("worst",)):
            _Eng.calls.append((int(s0), int(cnt_)))
            tot = np.zeros((x2.shape[0], 4))
            for sid in range(s0, s0 + cnt_):
                w = o.sample_gaussian(mean, std, o.split_flat(L, o.philox_normals(seed, sid, o.param_count(L))), order="f32")
                lo, hi = o.ibp_forward(L, w, x2, e)
                worst = o.worst_case_logits(lo, hi, labels)
                ex = np.exp(worst - worst.max(1, keepdims=True))
                tot += ex / ex.sum(1, keepdims=True)
            return {"worst": torch.from_numpy(tot.astype(np.float32))}

    class _Model:
        seed, shard, mode = 5, True, "fp32"
        _next = 100

        def __init__(self):
            self._e = _Eng()

        def engine(self):
            return self._e

        def _take_ids(self, k):
            v = _Model._next
            _Model._next += k
            return v

    mdl = _Model()
    cls = np.array([0, 1, 2, 3, 0])
    got = verifiers.chernoff_bound_verification(mdl, x, 0.01, cls, confidence=0.6, delta=0.3)       # n = 6 samples
    nver = verifiers.chernoff_sample_bound(0.6, 0.3)
    mdl2 = _Model(); mdl2.shard = False; _Model._next = 100
    full = verifiers.chernoff_bound_verification(mdl2, x, 0.01, cls, confidence=0.6, delta=0.3)
    offv, cntv = dd.shard_range(nver, rank, world)
    ok = ok and np.allclose(got, full, rtol=1e-5, atol=1e-6) and _Eng.calls[0] == (100 + offv, cntv) and _Eng.calls[1] == (100, nver)
    # sharded upload of the shared input batch: every rank contributes ceil(B / world) rows, one all-gather completes it (even and
    # ragged row counts); the "upload" of this CPU stand-in is a host copy
    for Bx in (6, 5):
        xs = o.synthetic_x((Bx, 11), seed=3)
        seen = []

        def up(a):
            seen.append(a.shape[0])
            return torch.from_numpy(np.ascontiguousarray(a))
        full = dd.sharded_upload(xs, up, torch.device("cpu"))
        per = (Bx + world - 1) // world
        ok = ok and full.shape == (Bx, 11) and np.array_equal(full.numpy(), xs) and seen == [min(per, max(Bx - rank * per, 0))]
    ok = ok and dd.sharded_upload(o.synthetic_x((1, 11), seed=3), up, torch.device("cpu")) is None      # fewer rows than ranks
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
