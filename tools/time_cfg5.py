"""Config 5 (count_predicate, 128 inputs) timing at a few sample counts; env toggles are read by the library."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from tools.bench_configs import mlp_engine, timeit
from oracle import oracle as o
eng = mlp_engine([784, 512, 512, 10])
x = torch.from_numpy(o.synthetic_x((128, 784), seed=0)).cuda()
lab = np.random.Generator(np.random.Philox(3)).integers(0, 10, size=128).astype(np.int32)
n = int(os.environ.get("N", "20000"))
ms = timeit(lambda: eng.count_predicate(x, lab, n, seed=5, mode="bf16"), warm=1, reps=3)
print({k: v for k, v in os.environ.items() if k.startswith("DBX_")}, "n", n, "ms", round(ms, 3), "us/256-chunk", round(ms * 1000 / (n / 256), 1), eng.last_path())
