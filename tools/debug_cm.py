import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import deepbayes_b200 as db
from deepbayes_b200.engine import Engine
from oracle import oracle as o

def run(dims, B, n, cm, ns):
    L = o.mlp(dims)
    layers = [db.Dense(dims[1], activation="relu", input_shape=(dims[0],))]
    for h in dims[2:-1]:
        layers.append(db.Dense(h, activation="relu"))
    layers.append(db.Dense(dims[-1], activation="softmax"))
    eng = Engine(db.Sequential(layers))
    mean, std = o.synthetic_posterior(L, seed=5, sigma_scale=0.5)
    eng.set_gaussian(mean, std)
    x = o.synthetic_x((B, dims[0]), seed=6)
    os.environ["DBX_NO_FUSED"] = "1"
    g1, _ = eng.predict_sums(x, n, seed=17, s0=3, mode="bf16")
    os.environ["DBX_NO_FUSED"] = "0"
    os.environ["DBX_FUSED_CM"] = str(cm)
    os.environ["DBX_FUSED_NS"] = str(ns)
    f1, _ = eng.predict_sums(x, n, seed=17, s0=3, mode="bf16")
    f = (f1 / n).cpu().numpy(); g = (g1 / n).cpu().numpy()
    bad = ~np.isfinite(f).all(1) | (np.abs(f - g).max(1) > 2e-3)
    rows = np.nonzero(bad)[0]
    # does a bad tile match the generic result of some OTHER tile (X mix-up)?
    T = (B + 127) // 128
    match = {}
    for t in sorted(set((rows // 128).tolist())):
        lo, hi = t * 128, min(B, t * 128 + 128)
        for t2 in range(T):
            lo2, hi2 = t2 * 128, min(B, t2 * 128 + 128)
            if hi2 - lo2 == hi - lo and np.isfinite(f[lo:hi]).all() and np.abs(f[lo:hi] - g[lo2:hi2]).max() < 2e-3:
                match[t] = t2
    print("   tile->matching generic tile:", match)
    print("dims", dims, "B", B, "n", n, "CM", cm, "ns", ns, "-> bad rows:", len(rows),
          "tiles:", sorted(set((rows // 128).tolist())), "nan rows:", int((~np.isfinite(f).all(1)).sum()),
          "first:", rows[:8].tolist(), "maxdiff(finite):", float(np.nanmax(np.abs(f - g))))

if __name__ == "__main__":
    d2 = [784, 512, 512, 10]
    os.environ["DBX_FUSED_KERNEL"] = sys.argv[1] if len(sys.argv) > 1 else "3"
    run(d2, 256, 1, 2, 16)
    run(d2, 1024, 16, 2, 16)
    run(d2, 1000, 70, 2, 16)
    run([784, 512, 10], 700, 33, 2, 8)
    run([784, 128, 10], 300, 20, 2, 8)
    run([100, 256, 256, 7], 513, 40, 2, 16)
    run([784, 384, 384, 10], 260, 12, 2, 16)
    run(d2, 4096, 256, 2, 128)
