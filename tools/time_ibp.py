"""Time one chunk-sized IBP call (bf16) with CUDA events; prints ms per call."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from tools.bench_configs import mlp_engine
from oracle import oracle as o
eng = mlp_engine([784, 512, 512, 10])
x = torch.from_numpy(o.synthetic_x((128, 784), seed=0)).cuda()
lab = np.random.Generator(np.random.Philox(3)).integers(0, 10, size=128).astype(np.int32)
n = int(os.environ.get("N", "2560"))
for _ in range(3):
    eng.ibp_predict(x, 0.01, lab, n, seed=5, mode="bf16", want=("worst", "counts"))
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    eng.ibp_predict(x, 0.01, lab, n, seed=5, mode="bf16", want=("worst", "counts"))
b.record(); torch.cuda.synchronize()
print("DBG", os.environ.get("DBX_IBP_DBG", "0"), "ms/call", a.elapsed_time(b) / 5, "us/chunk", a.elapsed_time(b) / 5 * 1000 / (n / 256))
