"""A few bf16 training steps of MLP 784-512-512-10 at one mini-batch size, for a launch list:
   ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/train_launches.csv python tools/time_train_step.py 1024"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import deepbayes_b200 as db
from tools.bench_configs import o

bs = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
layers = [db.Dense(512, activation="relu", input_shape=(784,)), db.Dense(512, activation="relu"), db.Dense(10, activation="softmax")]
opt = db.optimizers.BayesByBackprop().compile(db.Sequential(layers), None, batch_size=bs, inflate_prior=1.0, seed=3, train_mode="bf16")
x = torch.from_numpy(o.synthetic_x((bs, 784), seed=0)).cuda()
y = np.random.Generator(np.random.Philox(3)).integers(0, 10, size=bs)
for _ in range(4):
    opt.step(x, y, 0.01, sync=False)
torch.cuda.synchronize()
print(opt.engine().last_path())
