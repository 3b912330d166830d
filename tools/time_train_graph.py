"""Training step inside dbx_bbb_train_epoch (bf16 mode): CUDA-graph replay against step-by-step launches, per mini-batch size and
epoch length.  python tools/time_train_graph.py"""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import deepbayes_b200 as db
from deepbayes_b200 import _lib
from tools.bench_configs import timeit, o

dims = [784, 512, 512, 10]
for bs, nb in ((128, 40), (128, 400), (1024, 40), (1024, 200)):
    row = {"config": "dbx_bbb_train_epoch bf16, MLP 784-512-512-10", "B": bs, "steps_per_call": nb}
    for no_graph in (1, 0):
        _lib.set_option("TRAIN_GRAPH", 0 if no_graph else 1)
        layers = [db.Dense(512, activation="relu", input_shape=(784,)), db.Dense(512, activation="relu"), db.Dense(10, activation="softmax")]
        opt = db.optimizers.BayesByBackprop().compile(db.Sequential(layers), None, batch_size=bs, inflate_prior=1.0, seed=3, train_mode="bf16")
        x = torch.from_numpy(o.synthetic_x((bs, 784), seed=0)).cuda()
        y = np.random.Generator(np.random.Philox(3)).integers(0, 10, size=bs)
        opt.step(x, y, 0.01, sync=False)
        eng = opt.engine()
        N = bs * nb
        Xd = torch.from_numpy(o.synthetic_x((N, 784), seed=1)).cuda()
        yd = torch.from_numpy(np.random.Generator(np.random.Philox(4)).integers(0, 10, size=N).astype(np.int32)).cuda()
        ep = lambda: eng.bbb_train_epoch(Xd, yd, bs, opt._d_mean, opt._d_rho, opt._d_pm, opt._d_pv, 0.01, 1.0, 3, 0, want_weights=False, mode="bf16")
        import time
        ep(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            ep()
        torch.cuda.synchronize()
        row["ms_per_step_" + ("steps" if no_graph else "graph")] = (time.perf_counter() - t0) / 3 / nb * 1e3
    print(json.dumps(row), flush=True)
_lib.reset_options()
