import os, sys
sys.path.insert(0, "/root/repo")
import torch, numpy as np
from tools.bench_configs import mlp_engine, timeit
from oracle import oracle as o
S = 512
eng = mlp_engine([768, 1024, 1024, 10])
slab = torch.empty((S, eng.P), dtype=torch.float32, device="cuda")
for s0 in range(0, S, 256):
    slab[s0:s0 + 256] = eng.sample(256, seed=7, s0=s0)
eng.set_samples(slab)
x = torch.from_numpy(o.synthetic_x((64, 768), seed=0)).cuda()
for _ in range(3):
    eng.predict_stored_sums(x, S, j0=0, mode="bf16")
torch.cuda.synchronize()
print(eng.last_path())
