import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import deepbayes_b200 as db
from deepbayes_b200.engine import Engine
from oracle import oracle as o
from tools.bench_configs import timeit
m = db.Sequential([db.Conv2D(32, 3, activation="relu", input_shape=(32, 32, 3)), db.Conv2D(64, 3, activation="relu"),
                   db.MaxPooling2D(2), db.Flatten(), db.Dense(128, activation="relu"), db.Dense(10, activation="softmax")])
eng = Engine(m)
mean, std = o.synthetic_posterior(o.cnn_cfg4(), seed=1)
eng.set_gaussian(mean, std)
B, n = 256, int(sys.argv[1]) if len(sys.argv) > 1 else 16
x = torch.from_numpy(o.synthetic_x((B, 32, 32, 3), seed=0)).cuda()
ms = timeit(lambda: eng.predict_sums(x, n, seed=1, mode="bf16"), warm=1, reps=2)
print("cfg4 n", n, "ms", ms, "ms/sample", ms / n, eng.last_path())
