"""Config 3 at one batch size, with / without the fp32-slab tensor-core path (A/B for the routing threshold)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tools.bench_configs import mlp_engine, timeit
from oracle import oracle as o
B = int(sys.argv[1]); S = 1000
eng = mlp_engine([784, 1024, 1024, 10])
slab = torch.empty((S, eng.P), dtype=torch.float32, device="cuda")
for s0 in range(0, S, 250):
    slab[s0:s0 + 250] = eng.sample(250, seed=7, s0=s0)
eng.set_samples(slab)
x = torch.from_numpy(o.synthetic_x((B, 784), seed=0)).cuda()
ms = timeit(lambda: eng.predict_stored_sums(x, S, j0=0, mode="bf16"), warm=1, reps=3)
print("B", B, "NO_F32W", os.environ.get("DBX_NO_F32W", "0"), "ms", round(ms, 3), "GB/s", round(4.0 * eng.P * S / (ms * 1e-3) / 1e9), eng.last_path())
