"""Raw per-warp event dump of the fused-kernel timeline for one steady-state unit (debug aid)."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["DBX_FUSED_TRACE"] = "1"
import deepbayes_b200 as db
from deepbayes_b200 import _lib
from deepbayes_b200.engine import Engine
from oracle import oracle as o

L = o.mlp([784, 512, 512, 10])
m = db.Sequential([db.Dense(512, activation="relu", input_shape=(784,)), db.Dense(512, activation="relu"), db.Dense(10, activation="softmax")])
eng = Engine(m)
mean, std = o.synthetic_posterior(L, seed=1)
eng.set_gaussian(mean, std)
x = o.synthetic_x((4096, 784), seed=0)
for _ in range(2):
    eng.predict_sums(x, 256, seed=1, mode="bf16")
lib = _lib.load()
buf = (C.c_longlong * 16000)()
lib.dbx_debug_trace.restype = C.c_int
n = lib.dbx_debug_trace(buf, 8000)
ev = np.array(buf[:2 * n]).reshape(n, 2)
ev = ev[np.argsort(ev[:, 1], kind="stable")]
starts = [i for i in range(n) if ev[i, 0] == 100]
a, b = starts[4], starts[5]
lo, hi = int(sys.argv[1]) if len(sys.argv) > 1 else 0, int(sys.argv[2]) if len(sys.argv) > 2 else 120
for i in range(a + lo, min(a + hi, b)):
    print("%8d +%5d  %d" % (ev[i, 1] - ev[a, 1], ev[i, 1] - ev[i - 1, 1], ev[i, 0]))
