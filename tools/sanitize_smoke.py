"""Small end-to-end run of every kernel family for compute-sanitizer (memcheck)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import deepbayes_b200 as db
from deepbayes_b200.engine import Engine
from oracle import oracle as o

def mlp(dims):
    layers = [db.Dense(dims[1], activation="relu", input_shape=(dims[0],))]
    for h in dims[2:-1]:
        layers.append(db.Dense(h, activation="relu"))
    layers.append(db.Dense(dims[-1], activation="softmax"))
    eng = Engine(db.Sequential(layers))
    mean, std = o.synthetic_posterior(o.mlp(dims), seed=5, sigma_scale=0.5)
    eng.set_gaussian(mean, std)
    return eng

for dims, B, n in (([784, 512, 512, 10], 300, 5), ([784, 256, 10], 130, 3), ([784, 384, 384, 10], 260, 3)):
    eng = mlp(dims)
    x = o.synthetic_x((B, dims[0]), seed=6)
    for kern in (1, 2):
        db._lib.set_option("FUSED_KERNEL", kern)
        s1, s2 = eng.predict_sums(x, n, seed=1, mode="bf16", second_moment=True)
        print(dims, "kernel", kern, eng.last_path(), float(s1.sum()))
    db._lib.set_option("FUSED_KERNEL", 0)
    db._lib.set_option("NO_FUSED", 1)
    s1, _ = eng.predict_sums(x, n, seed=1, mode="bf16"); print("generic_tc", eng.last_path(), float(s1.sum()))
    s1, _ = eng.predict_sums(x, n, seed=1, mode="fp32"); print("generic fp32", eng.last_path(), float(s1.sum()))
    db._lib.set_option("NO_FUSED", 0)
    s1, _ = eng.predict_sums(x[:7], n, seed=1, mode="bf16"); print("stream", eng.last_path(), float(s1.sum()))
    lab = np.zeros(B, np.int32)
    c = eng.count_predicate(x, lab, n, seed=1, mode="bf16"); print("count", eng.last_path(), int(c.sum()))
m = db.Sequential([db.Conv2D(8, 3, activation="relu", input_shape=(14, 14, 8)), db.Conv2D(16, 3, activation="relu", padding="same", strides=2),
                   db.MaxPooling2D(2), db.Flatten(), db.Dense(24, activation="relu"), db.Dense(10, activation="softmax")])
eng = Engine(m)
L = [o.Conv2D(3, 3, 8, 8, "relu"), o.Conv2D(3, 3, 8, 16, "relu", (2, 2), "same"), o.MaxPool2D((2, 2)), o.Flatten(), o.Dense(3 * 3 * 16, 24, "relu"), o.Dense(24, 10, "softmax")]
mean, std = o.synthetic_posterior(L, seed=1)
eng.set_gaussian(mean, std)
x = o.synthetic_x((5, 14, 14, 8), seed=2)
for mode in ("bf16", "fp32"):
    s1, _ = eng.predict_sums(x, 3, seed=1, mode=mode); print("cnn", mode, eng.last_path(), float(s1.sum()))
want = o.predict(L, mean, std, x, 3, [[e for e in s] for s in o.synthetic_eps(L, 3, seed=9)])
eps = np.stack([o.flatten_list(e) for e in o.synthetic_eps(L, 3, seed=9)])
got, _ = eng.predict_sums(x, 3, eps=eps, mode="bf16")
print("cnn tc vs oracle max|d| =", float(np.abs((got / 3).cpu().numpy() - want).max()))
print("done")
