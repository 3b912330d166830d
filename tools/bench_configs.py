"""Measure the other BASELINE.json configs (1, 3, 4, 5) on one B200 and print one JSON line each.
Usage: python tools/bench_configs.py [--configs 1,3,4,5] [--stored 2000]"""
import argparse, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import deepbayes_b200 as db
from deepbayes_b200 import _lib
from deepbayes_b200.engine import Engine
from oracle import oracle as o

PEAKS = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))) \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else \
    {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}


def timeit(fn, warm=2, reps=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def mlp_engine(dims):
    layers = [db.Dense(dims[1], activation="relu", input_shape=(dims[0],))]
    for h in dims[2:-1]:
        layers.append(db.Dense(h, activation="relu"))
    layers.append(db.Dense(dims[-1], activation="softmax"))
    eng = Engine(db.Sequential(layers))
    mean, std = o.synthetic_posterior(o.mlp(dims), seed=1)
    eng.set_gaussian(mean, std)
    return eng


def emit(**kw):
    print(json.dumps(kw), flush=True)


def cfg1():
    eng = mlp_engine([784, 512, 10])
    for B in (10000, 128, 1):
        x = torch.from_numpy(o.synthetic_x((B, 784), seed=0)).cuda()
        n = 100
        ms = timeit(lambda: eng.predict_sums(x, n, seed=1, mode="bf16", second_moment=True))
        fl = eng.flops_per_forward * B * n / (ms * 1e-3) / 1e12
        emit(config="1: BBB MLP 784-512-10, n=100", B=B, ms=ms, value=B * n / (ms * 1e-3), unit="forwards/s", path=eng.last_path(),
             tflops=fl, frac_of_bf16_burst=fl / PEAKS["bf16_tflops"])


def cfg3(S):
    eng = mlp_engine([784, 1024, 1024, 10])
    P = eng.P
    slab = torch.empty((S, P), dtype=torch.float32, device="cuda")
    for s0 in range(0, S, 256):   # synthetic stored samples = Gaussian draws, written straight into the slab
        nc = min(256, S - s0)
        slab[s0:s0 + nc] = eng.sample(nc, seed=7, s0=s0)
    eng.set_samples(slab)
    freq = torch.full((S,), 1.0 / S, device="cuda")
    for B in (1, 8, 16, 64, 128, 512):
        x = torch.from_numpy(o.synthetic_x((B, 784), seed=0)).cuda()
        lib = _lib.load()
        ms = timeit(lambda: eng.predict_stored_sums(x, S, j0=0, wts=freq, mode="bf16", second_moment=True), warm=1, reps=3)
        gbs = 4.0 * P * S / (ms * 1e-3) / 1e9
        emit(config="3: HMC stored posterior, MLP 784-1024-1024-10, %d stored fp32 samples resident in HBM (%.1f GB)" % (S, 4e-9 * P * S),
             B=B, ms=ms, value=S * B / (ms * 1e-3), unit="forwards/s", samples_per_s=S / (ms * 1e-3), path=eng.last_path(),
             roofline={"bound": "hbm", "achieved": gbs, "peak": PEAKS["hbm_gbs"], "unit": "GB/s", "frac": gbs / PEAKS["hbm_gbs"],
                       "algorithmic_bytes_per_sample": 4 * P})
    del slab


def cfg4():
    m = db.Sequential([db.Conv2D(32, 3, activation="relu", input_shape=(32, 32, 3)), db.Conv2D(64, 3, activation="relu"),
                       db.MaxPooling2D(2), db.Flatten(), db.Dense(128, activation="relu"), db.Dense(10, activation="softmax")])
    eng = Engine(m)
    mean, std = o.synthetic_posterior(o.cnn_cfg4(), seed=1)
    eng.set_gaussian(mean, std)
    B, n = 256, 256
    x = torch.from_numpy(o.synthetic_x((B, 32, 32, 3), seed=0)).cuda()
    ms = timeit(lambda: eng.predict_sums(x, n, seed=1, mode="bf16"), warm=1, reps=2)
    fl = eng.flops_per_forward * B * n / (ms * 1e-3) / 1e12
    emit(config="4: Bayesian CNN conv32-conv64-pool-dense128-dense10, CIFAR-10 shape, n=256", B=B, ms=ms,
         value=B * n / (ms * 1e-3), unit="forwards/s", path=eng.last_path(), tflops=fl, frac_of_bf16_burst=fl / PEAKS["bf16_tflops"],
         note="generic bf16 path: first conv (Cin = 3) = im2col of the shared input (once per call) + dense_tc_kernel; second conv = conv_tc_kernel (direct, per-tile input patch by TMA); Dense = dense_tc_kernel")


def cfg5():
    eng = mlp_engine([784, 512, 512, 10])
    K, n = 128, 100000
    g = np.random.Generator(np.random.Philox(3))
    x0 = o.synthetic_x((K, 784), seed=0)
    xadv = np.clip(x0 + g.uniform(-0.1, 0.1, x0.shape).astype(np.float32), 0, 1)
    x = torch.from_numpy(xadv).cuda()
    lab = g.integers(0, 10, size=K).astype(np.int32)
    ms = timeit(lambda: eng.count_predicate(x, lab, n, seed=5, mode="bf16"), warm=1, reps=2)
    fl = eng.flops_per_forward * K * n / (ms * 1e-3) / 1e12
    emit(config="5: Chernoff-bound MC, 100k posterior samples x 128 perturbed inputs, MLP 784-512-512-10", B=K, ms=ms,
         value=K * n / (ms * 1e-3), unit="forwards/s", samples_per_s=n / (ms * 1e-3), path=eng.last_path(), tflops=fl,
         frac_of_bf16_burst=fl / PEAKS["bf16_tflops"])
    for K1 in (1,):
        ms = timeit(lambda: eng.count_predicate(x[:K1], lab[:K1], n, seed=5, mode="bf16"), warm=1, reps=2)
        emit(config="5b: massart-style loop, 100k posterior samples x 1 input", B=K1, ms=ms, samples_per_s=n / (ms * 1e-3),
             path=eng.last_path(), note="weights formed in registers from Philox, never materialised")
    # 5c: the same loop with verify=True -- IBP over the eps-ball for every posterior sample
    n2 = 10000
    for mode in ("bf16", "fp32"):
        ms = timeit(lambda: eng.ibp_predict(x, 0.01, lab, n2, seed=5, mode=mode, want=("worst", "counts")), warm=1, reps=2)
        fl = 2 * 1337344 * K * n2 / (ms * 1e-3) / 1e12          # centre and radius products
        emit(config="5c: Chernoff-bound MC with IBP (verify=True), 10k posterior samples x 128 eps-balls, MLP 784-512-512-10",
             B=K, mode=mode, ms=ms, value=K * n2 / (ms * 1e-3), unit="interval forwards/s", samples_per_s=n2 / (ms * 1e-3),
             path=eng.last_path(), tflops=fl)


def cfg_train():
    """BBB training throughput: MLP 784-512-10 and 784-512-512-10, mini-batch 128, one `dbx_bbb_step` per step
    with the parameters resident on the device (sync=False)."""
    import deepbayes_b200 as db
    for dims in ([784, 512, 10], [784, 512, 512, 10]):
        layers = [db.Dense(dims[1], activation="relu", input_shape=(dims[0],))]
        for h in dims[2:-1]:
            layers.append(db.Dense(h, activation="relu"))
        layers.append(db.Dense(dims[-1], activation="softmax"))
        opt = db.optimizers.BayesByBackprop().compile(db.Sequential(layers), None, batch_size=128, inflate_prior=1.0, seed=3)
        x = torch.from_numpy(o.synthetic_x((128, 784), seed=0)).cuda()
        y = np.random.Generator(np.random.Philox(3)).integers(0, 10, size=128)
        opt.step(x, y, 0.01, sync=False)

        def run():
            for _ in range(50):
                opt.step(x, y, 0.01, sync=False)
        ms = timeit(run, warm=1, reps=3) / 50
        emit(config="train: BayesByBackprop.step, MLP %s, mini-batch 128, fp32" % "-".join(map(str, dims)), B=128, ms=ms,
             value=1e3 / ms, unit="steps/s", path="dbx_bbb_step", note="reference: 8.4 ms per sample-forward on CPU (SURVEY section 6); one step = sample + forward + backward + update")
        # the same steps as ONE device call per epoch (dbx_bbb_train_epoch): 200 mini-batches of 128 resident examples
        eng = opt.engine()
        N = 128 * 200
        Xd = torch.from_numpy(o.synthetic_x((N, 784), seed=1)).cuda()
        yd = torch.from_numpy(np.random.Generator(np.random.Philox(4)).integers(0, 10, size=N).astype(np.int32)).cuda()
        ep = lambda: eng.bbb_train_epoch(Xd, yd, 128, opt._d_mean, opt._d_rho, opt._d_pm, opt._d_pv, 0.01, 1.0, 3, 0, want_weights=False)
        ms2 = timeit(ep, warm=1, reps=3) / 200
        emit(config="train: one dbx_bbb_train_epoch call per 200 mini-batches, MLP %s, mini-batch 128, fp32" % "-".join(map(str, dims)), B=128,
             ms=ms2, value=1e3 / ms2, unit="steps/s", path="dbx_bbb_train_epoch")


def cfg_train_tc():
    """The training step per mini-batch size, fp32 kernels vs tensor-core GEMMs (train_mode): one dbx_bbb_train_epoch call over 40
    mini-batches of resident examples, MLP 784-512-512-10."""
    import deepbayes_b200 as db
    dims = [784, 512, 512, 10]
    for bs in (128, 256, 1024, 4096):
        for tm in ("fp32", "bf16"):
            layers = [db.Dense(dims[1], activation="relu", input_shape=(dims[0],))]
            for h in dims[2:-1]:
                layers.append(db.Dense(h, activation="relu"))
            layers.append(db.Dense(dims[-1], activation="softmax"))
            opt = db.optimizers.BayesByBackprop().compile(db.Sequential(layers), None, batch_size=bs, inflate_prior=1.0, seed=3, train_mode=tm)
            x = torch.from_numpy(o.synthetic_x((bs, 784), seed=0)).cuda()
            y = np.random.Generator(np.random.Philox(3)).integers(0, 10, size=bs)
            opt.step(x, y, 0.01, sync=False)
            eng = opt.engine()
            nb = 40 if bs <= 1024 else 10
            N = bs * nb
            Xd = torch.from_numpy(o.synthetic_x((N, 784), seed=1)).cuda()
            yd = torch.from_numpy(np.random.Generator(np.random.Philox(4)).integers(0, 10, size=N).astype(np.int32)).cuda()
            ep = lambda: eng.bbb_train_epoch(Xd, yd, bs, opt._d_mean, opt._d_rho, opt._d_pm, opt._d_pv, 0.01, 1.0, 3, 0, want_weights=False, mode=tm)
            ms = timeit(ep, warm=1, reps=3) / nb
            emit(config="train: dbx_bbb_train_epoch, MLP 784-512-512-10, mini-batch %d" % bs, B=bs, mode=tm, ms=ms, value=bs * 1e3 / ms,
                 unit="examples/s", path="dbx_bbb_train_epoch/" + eng.last_path(), tflops=3 * 1337344 * bs / (ms * 1e-3) / 1e12)


def cfg_attack():
    """Expected input gradient: MLP 784-512-512-10, 128 inputs, one FGSM gradient = 10 iterations x 35
    posterior samples (a PosteriorModel's predict), all on the device; and config 5 with verify=False (massart_bound_check's
    FGSM per posterior sample, verifiers.py:622-634): 10,000 samples x 128 inputs.  fp32 kernels and the tensor-core path."""
    eng = mlp_engine([784, 512, 512, 10])
    x = torch.from_numpy(o.synthetic_x((128, 784), seed=0)).cuda()
    y = np.random.Generator(np.random.Philox(3)).integers(0, 10, size=128)
    for mode in ("fp32", "bf16"):
        ms = timeit(lambda: eng.input_gradient(x, y, 10, 35, seed=5, mode=mode), warm=1, reps=3)
        fl = 2 * 1337344 * 128 * 350 / (ms * 1e-3) / 1e12     # forward + backward-data
        emit(config="attack: gradient_expectation, MLP 784-512-512-10, 128 inputs, num_models=10 x 35 samples", B=128, ms=ms, mode=mode,
             value=350 * 128 / (ms * 1e-3), unit="sample-input gradients/s", path="dbx_input_gradient/" + eng.last_path(), tflops=fl)
    for mode, n in (("fp32", 2000), ("bf16", 10000)):
        ms = timeit(lambda: eng.count_predicate_fgsm(x, y, y, 0.05, n, seed=5, lower=0.0, upper=1.0, mode=mode), warm=1, reps=2)
        fl = 3 * 1337344 * 128 * n / (ms * 1e-3) / 1e12       # forward + backward-data + forward at the adversarial input
        emit(config="config 5 with verify=False: FGSM per posterior sample, MLP 784-512-512-10, %d samples x 128 inputs" % n, B=128, n=n,
             ms=ms, mode=mode, value=n / (ms * 1e-3), unit="posterior samples/s", path="dbx_count_predicate_fgsm/" + eng.last_path(), tflops=fl)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="1,3,4,5")
    ap.add_argument("--stored", type=int, default=2000)
    a = ap.parse_args()
    for c in a.configs.split(","):
        {"1": cfg1, "3": lambda: cfg3(a.stored), "4": cfg4, "5": cfg5, "train": cfg_train, "train_tc": cfg_train_tc, "attack": cfg_attack}[c]()
