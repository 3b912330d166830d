"""IBP bf16 path, one chunk-sized call (for ncu launch lists)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from tools.bench_configs import mlp_engine
from oracle import oracle as o
eng = mlp_engine([784, 512, 512, 10])
x = torch.from_numpy(o.synthetic_x((128, 784), seed=0)).cuda()
lab = np.random.Generator(np.random.Philox(3)).integers(0, 10, size=128).astype(np.int32)
for _ in range(2):
    eng.ibp_predict(x, 0.01, lab, 512, seed=5, mode="bf16", want=("worst", "counts"))
torch.cuda.synchronize()
print(eng.last_path())
