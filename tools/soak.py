"""Soak: random shapes through the newest persistent kernels (fused IBP layers, conv + fused max-pool, fused MLP) for a
fixed wall time; every call is followed by a synchronize, so a hang shows up as a timeout of the whole script."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import deepbayes_b200 as db
from deepbayes_b200.engine import Engine
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
g = np.random.Generator(np.random.Philox(99))
t0 = time.time(); it = 0
while time.time() - t0 < budget:
    kind = it % 6
    if os.environ.get('SOAK_VERBOSE'):
        print('it', it, 'kind', kind, 'start', flush=True)
    if kind == 0:      # MLP: IBP (fused layers) + predict (fused / generic)
        K = int(g.integers(1, 40)) * 8
        hs = [int(g.integers(1, 80)) * 8 for _ in range(int(g.integers(1, 4)))]
        C = int(g.integers(2, 17))
        layers = [db.Dense(hs[0], activation="relu", input_shape=(K,))] + [db.Dense(h, activation=str(g.choice(["relu", "tanh"]))) for h in hs[1:]]
        layers.append(db.Dense(C, activation="softmax"))
        eng = Engine(db.Sequential(layers))
        eng.set_gaussian((g.standard_normal(eng.P) * 0.1).astype(np.float32), np.full(eng.P, 0.02, np.float32))
        B, n = int(g.integers(1, 700)), int(g.integers(1, 600))
        x = g.random((B, K)).astype(np.float32)
        lab = g.integers(0, C, size=B).astype(np.int32)
        out = eng.ibp_predict(x, 0.01, lab, n, seed=it, mode="bf16", want=("worst", "lower", "upper", "counts"))
        assert bool((out["lower"] <= out["upper"]).all())
        s1, _ = eng.predict_sums(x, n, seed=it, mode="bf16")
        assert bool(torch.isfinite(s1).all())
    elif kind == 1:    # CNN with a fusable pool
        H, W = int(g.integers(8, 34)), int(g.integers(8, 34))
        c1, c2 = int(g.choice([16, 32, 64])), int(g.choice([8, 16, 32, 64]))
        m = db.Sequential([db.Conv2D(c1, 3, activation="relu", input_shape=(H, W, 3)), db.Conv2D(c2, 3, padding=str(g.choice(["valid", "same"])), activation="relu"),
                           db.MaxPooling2D(2), db.Flatten(), db.Dense(16, activation="relu"), db.Dense(10, activation="softmax")])
        eng = Engine(m)
        eng.set_gaussian((g.standard_normal(eng.P) * 0.08).astype(np.float32), np.full(eng.P, 0.01, np.float32))
        B, n = int(g.integers(1, 80)), int(g.integers(1, 60))
        x = g.random((B, H, W, 3)).astype(np.float32)
        s1, _ = eng.predict_sums(x, n, seed=it, mode="bf16")
        assert bool(torch.isfinite(s1).all())
    elif kind == 3:    # stored posterior: fp32 slab through dense_tc_f32w_kernel (64- and 128-column tiles, 64- / 128-row A boxes)
        K = int(g.integers(2, 30)) * 8
        hs = [int(g.choice([128, 192, 256, 320, 512, 1024])) for _ in range(int(g.integers(1, 3)))]
        layers = [db.Dense(hs[0], activation="relu", input_shape=(K,))] + [db.Dense(h, activation="relu") for h in hs[1:]] + [db.Dense(10, activation="softmax")]
        eng = Engine(db.Sequential(layers))
        eng.set_gaussian((g.standard_normal(eng.P) * 0.1).astype(np.float32), np.full(eng.P, 0.02, np.float32))
        S = int(g.integers(2, 40))
        eng.set_samples(eng.sample(S, seed=it, s0=0))
        B = int(g.integers(9, 300))
        x = g.random((B, K)).astype(np.float32)
        a, _ = eng.predict_stored_sums(x, S, j0=0, mode="bf16")
        b, _ = eng.predict_stored_sums(x, S, j0=0, mode="fp32")
        assert float((a - b).abs().max()) / S < 2e-2
    elif kind == 4:    # gradient loops on the tensor cores: dense_tc_bwd_kernel (persistent / per-tile), FGSM epilogue
        K = int(g.integers(1, 110)) * 8
        hs = [int(g.integers(1, 80)) * 8 for _ in range(int(g.integers(1, 4)))]
        C = int(g.integers(2, 17))
        layers = [db.Dense(hs[0], activation="relu", input_shape=(K,))] + [db.Dense(h, activation=str(g.choice(["relu", "tanh", "sigmoid"]))) for h in hs[1:]]
        layers.append(db.Dense(C, activation="softmax"))
        eng = Engine(db.Sequential(layers))
        eng.set_gaussian((g.standard_normal(eng.P) * 0.1).astype(np.float32), np.full(eng.P, 0.02, np.float32))
        B, n = int(g.integers(1, 300)), int(g.integers(1, 400))
        x = g.random((B, K)).astype(np.float32)
        lab = g.integers(0, C, size=B).astype(np.int32)
        gr = eng.input_gradient(x, lab, int(g.integers(1, 4)), min(n, 64), seed=it, mode="bf16")
        assert eng.last_path() == "generic_tc" and bool(torch.isfinite(gr).all())
        c = eng.count_predicate_fgsm(x, lab, lab, 0.02, n, seed=it, lower=0.0, upper=1.0, mode="bf16")
        assert int(c.max()) <= n
    elif kind == 5:    # training step on the tensor cores: dense_tc_wgrad_kernel, both robust modes
        K = int(g.integers(1, 110)) * 8
        hs = [int(g.integers(1, 80)) * 8 for _ in range(int(g.integers(1, 4)))]
        C = int(g.integers(2, 17))
        layers = [db.Dense(hs[0], activation="relu", input_shape=(K,))] + [db.Dense(h, activation=str(g.choice(["relu", "tanh"]))) for h in hs[1:]]
        layers.append(db.Dense(C, activation="softmax"))
        rt = int(g.choice([0, 2]))
        opt = db.optimizers.BayesByBackprop().compile(db.Sequential(layers), None, batch_size=64, inflate_prior=1.0, seed=it, train_mode="bf16",
                                                      robust_train=rt, epsilon=0.02, linear_schedule=False)
        B = int(g.integers(1, 700))
        x = g.random((B, K)).astype(np.float32)
        lab = g.integers(0, C, size=B).astype(np.int32)
        for _ in range(3):
            opt.step(x, lab, 0.05, sync=False)
        eng = opt.engine()
        assert eng.last_path() == "generic_tc"
        opt.sync_from_device()
        assert all(np.isfinite(np.asarray(t)).all() for t in opt.posterior_mean) and all(np.isfinite(np.asarray(t)).all() for t in opt.posterior_var)
        del opt
    else:              # fused MLP shapes, long runs at small batch
        H = int(g.choice([128, 256, 384, 512]))
        dims = [784, H, H, 10] if g.random() < 0.5 else [784, H, 10]
        layers = [db.Dense(dims[1], activation="relu", input_shape=(784,))] + [db.Dense(h, activation="relu") for h in dims[2:-1]] + [db.Dense(10, activation="softmax")]
        eng = Engine(db.Sequential(layers))
        eng.set_gaussian((g.standard_normal(eng.P) * 0.05).astype(np.float32), np.full(eng.P, 0.02, np.float32))
        B, n = int(g.integers(17, 3000)), int(g.integers(1, 3000))
        x = g.random((B, 784)).astype(np.float32)
        lab = g.integers(0, 10, size=B).astype(np.int32)
        c = eng.count_predicate(x, lab, n, seed=it, mode="bf16")
        assert int(c.max()) <= n
    torch.cuda.synchronize()
    if os.environ.get('SOAK_VERBOSE'):
        print('it', it, 'kind', kind, 'ok', round(time.time() - t0, 1), flush=True)
    del eng
    it += 1
    if it % 500 == 0:
        print('soak:', it, 'iterations', round(time.time() - t0, 1), 's', flush=True)
print("soak ok:", it, "iterations in", round(time.time() - t0, 1), "s")
