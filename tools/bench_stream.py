import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from tools.bench_configs import mlp_engine, timeit
from oracle import oracle as o
S = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
eng = mlp_engine([784, 1024, 1024, 10])
P = eng.P
slab = torch.empty((S, P), dtype=torch.float32, device="cuda")
for s0 in range(0, S, 256):
    nc = min(256, S - s0)
    slab[s0:s0 + nc] = eng.sample(nc, seed=7, s0=s0)
eng.set_samples(slab)
for B in (1, 8, 16):
    x = torch.from_numpy(o.synthetic_x((B, 784), seed=0)).cuda()
    ms = timeit(lambda: eng.predict_stored_sums(x, S, j0=0, mode="bf16"), warm=1, reps=3)
    print("B", B, "ms", ms, "GB/s", 4.0 * P * S / (ms * 1e-3) / 1e9, eng.last_path(), flush=True)
