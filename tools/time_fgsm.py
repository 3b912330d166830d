"""One bf16 dbx_count_predicate_fgsm call (config 5 with verify=False) for a launch list:
   ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/fgsm_launches.csv python tools/time_fgsm.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tools.bench_configs import mlp_engine, o

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
eng = mlp_engine([784, 512, 512, 10])
x = torch.from_numpy(o.synthetic_x((128, 784), seed=0)).cuda()
y = np.random.Generator(np.random.Philox(3)).integers(0, 10, size=128)
for _ in range(2):
    c = eng.count_predicate_fgsm(x, y, y, 0.05, n, seed=5, lower=0.0, upper=1.0, mode="bf16")
torch.cuda.synchronize()
print(eng.last_path(), c[:8].tolist())
