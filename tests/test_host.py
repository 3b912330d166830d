"""CPU tests of the host-side mirror of the reference interface (no CUDA calls)."""
import json
import math
import os

import numpy as np
import pytest

import deepbayes_b200 as db
from deepbayes_b200 import _lib, formats, distributed as dd
from deepbayes_b200.analyzers import verifiers


def test_library_loads_and_exports_every_declared_symbol():
    lib = _lib.load()
    names = _lib.declared_symbols()
    assert len(names) >= 20 and set(names) == set(_lib._SIGS)
    for n in names:
        assert hasattr(lib, n), n
    assert lib.dbx_version() >= 100


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    m = db.Sequential([db.Dense(4, activation="softmax", input_shape=(3,))])
    with pytest.raises(_lib.DbxError):
        m(np.zeros((2, 3), np.float32))


def test_sequential_json_roundtrip_and_keras_arch(golden_dir):
    m = db.Sequential([db.Conv2D(8, 3, activation="relu", input_shape=(14, 14, 3)), db.MaxPooling2D(2), db.Flatten(),
                       db.Dense(10, activation="softmax")])
    assert m.weight_shapes() == [(3, 3, 3, 8), (8,), (6 * 6 * 8, 10), (10,)]
    m2 = db.model_from_json(m.to_json())
    assert m2.weight_shapes() == m.weight_shapes() and [l.kind for l in m2.layers] == [l.kind for l in m.layers]
    with open(os.path.join(golden_dir, "posteriors", "VOGN_MNIST_Posterior", "arch.json")) as f:
        k = db.model_from_json(f.read())
    assert k.input_shape == (1, 784) and k.feature_shape == (784,)
    assert [l.activation.name for l in k.layers] == ["relu", "softmax"]
    rows, oshape = k.rows_and_outshape(np.zeros((5, 1, 784)))
    assert rows == 5 and oshape == (5, 1, 10)  # Appendix A.8
    ws = k.get_weights()
    k.set_weights([w + 1 for w in ws])
    np.testing.assert_array_equal(k.layers[1].get_weights()[1], ws[3] + 1)
    with pytest.raises(ValueError):
        k.set_weights(ws[:-1])


def test_compile_priors_match_oracle(oracle):
    L = oracle.cnn_cfg4()
    m = db.Sequential([db.Conv2D(32, 3, activation="relu", input_shape=(32, 32, 3)), db.Conv2D(64, 3, activation="relu"),
                       db.MaxPooling2D(2), db.Flatten(), db.Dense(128, activation="relu"), db.Dense(10, activation="softmax")])
    assert m.count_params() == 1626442
    opt = db.optimizers.BayesByBackprop().compile(m, None, inflate_prior=2.0, epochs=10)
    assert opt.epochs == 11  # linear_schedule default adds one epoch (optimizer.py:77-79)
    pm, pv = oracle.implicit_prior(L, 2.0)
    for a, b in zip(opt.prior_mean, pm):
        np.testing.assert_array_equal(a, b)
    for a, b in zip(opt.prior_var, pv):
        np.testing.assert_array_equal(a, b)
    # posterior mean starts at the prior mean (zeros), rho = log(exp(std)-1)  (bayesbybackprop.py:39-40)
    assert all((a == 0).all() for a in opt.posterior_mean)
    for r, v in zip(opt.posterior_var, pv):
        np.testing.assert_allclose(r, oracle.inverse_softplus(v), rtol=1e-6)
    # explicit scalar priors: element floor(i/2) per tensor (optimizer.py:250-256)
    opt2 = db.optimizers.StochasticGradientDescent().compile(m, None, prior_mean=0.5, prior_var=0.25)
    om, ov = oracle.prior_generator(L, 0.5, 0.25)
    for a, b in zip(opt2.prior_var, ov):
        np.testing.assert_array_equal(a, b)
    assert opt2.input_upper == math.inf and opt2.det is False and opt2.classes == 10


def test_formats_roundtrip(tmp_path, oracle):
    L = oracle.mlp([6, 5, 3])
    mean, std = oracle.synthetic_posterior(L, seed=1)
    m = db.Sequential([db.Dense(5, activation="relu", input_shape=(6,)), db.Dense(3, activation="softmax")])
    p = str(tmp_path / "g")
    formats.save_gaussian_posterior(p, m, mean, std, {"input_upper": 1.0, "input_lower": 0.0, "N": 7})
    assert formats.is_gaussian_dir(p)
    for a, b in zip(formats.load_tensor_list(p + "/mean.npy"), mean):
        np.testing.assert_array_equal(a, b)
    assert formats.load_info(p)["input_upper"] == 1.0
    assert formats.load_arch(p).weight_shapes() == m.weight_shapes()
    # HMC layout (hmc.py:308-322) + packed slab
    samples = [oracle.sample_gaussian(mean, std, e) for e in oracle.synthetic_eps(L, 4, seed=2)]
    q = str(tmp_path / "h")
    formats.save_sample_posterior(q, m, mean, samples, [3, 1, 2, 2])
    assert not formats.is_gaussian_dir(q) and formats.count_samples(q) == 4
    slab = formats.load_slab(q, oracle.param_count(L))
    np.testing.assert_array_equal(slab[2], oracle.flatten_list(samples[2]))
    os.remove(q + "/slab.npy")
    np.testing.assert_array_equal(formats.load_slab(q, oracle.param_count(L)), slab)
    assert formats.load_info(q)["input_lower"] == -math.inf  # Appendix A.2


def test_optimizer_save_formats(tmp_path, oracle):
    m = db.Sequential([db.Dense(5, activation="relu", input_shape=(6,)), db.Dense(3, activation="softmax")])
    bbb = db.optimizers.BayesByBackprop().compile(m, None)
    bbb.save(str(tmp_path / "bbb"))
    var = formats.load_tensor_list(str(tmp_path / "bbb" / "var.npy"))
    np.testing.assert_allclose(var[0], math.sqrt(1 / 6), rtol=1e-5)  # softplus(rho) round-trips to the prior std
    assert not os.path.exists(str(tmp_path / "bbb" / "info.pkl"))     # bayesbybackprop.py:184-195 writes none
    sgd = db.optimizers.StochasticGradientDescent().compile(m, None)
    sgd.save(str(tmp_path / "sgd"))
    info = formats.load_info(str(tmp_path / "sgd"))
    assert info["epochs"] == 11 and info["classes"] == 10 and "input_upper" in info
    hmc = db.optimizers.HamiltonianMonteCarlo().compile(m, None)
    for i in range(3):
        hmc.record([w + i for w in m.get_weights()], multiplicity=i + 1)
    hmc.save(str(tmp_path / "hmc"))
    np.testing.assert_array_equal(np.load(str(tmp_path / "hmc" / "freq.npy")), [1, 2, 3])
    assert formats.count_samples(str(tmp_path / "hmc")) == 3


def test_shard_range_partitions():
    for n in (1, 7, 1024, 100000):
        for G in (1, 2, 3, 8):
            parts = [dd.shard_range(n, r, G) for r in range(G)]
            assert sum(c for _, c in parts) == n
            assert parts[0][0] == 0 and all(parts[i][0] + parts[i][1] == parts[i + 1][0] for i in range(G - 1))
            assert max(c for _, c in parts) - min(c for _, c in parts) <= 1


def test_massart_helpers_match_oracle(oracle):
    assert verifiers.chernoff_sample_bound(0.95, 0.3) == oracle.chernoff_n(0.95, 0.3) == 380
    assert verifiers.okamoto_bound(0.05, 0.3) == oracle.okamoto_bound(0.05, 0.3)
    for succ, tr in ((5, 5), (0, 9), (40, 50), (3, 50)):
        I = list(verifiers.proportion_confint_beta(succ, tr))
        I = [0.0 if math.isnan(I[0]) else I[0], 1.0 if math.isnan(I[1]) else I[1]]
        assert tuple(I) == oracle.clopper_pearson(succ, tr)
        assert verifiers.absolute_massart_halting(succ, tr, I, 0.05, 0.3, 0.05) == \
            oracle.absolute_massart_halting(succ, tr, I, 0.05, 0.3, 0.05)


def test_bench_reference_arm_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the GPU arm): one JSON line with the contract's keys,
    on the default workload, also when torchrun has exported OMP_NUM_THREADS=1 (the arm restores the thread count)."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, DBX_REF_SAMPLES="1", OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "dtype", "config",
              "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["impl"] == "reference" and line["unit"] == "forwards/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and "784-512-512-10" in line["config"]["workload"]
    # ranks other than 0 print nothing and exit 0
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, env=dict(env, RANK="1", WORLD_SIZE="2"), timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_peer_window_api_refuses_bad_arguments():
    """The multi-GPU entry points fail with a message (not a crash) before any CUDA work when called wrongly."""
    import ctypes as C
    from deepbayes_b200 import _lib
    lib = _lib.load()
    h = C.c_void_p()
    assert lib.dbx_comm_create(3, 2, 0, 1024, C.byref(h)) != 0          # rank outside the world
    assert b"rank" in lib.dbx_last_error()
    assert lib.dbx_comm_create(0, 99, 0, 1024, C.byref(h)) != 0         # more ranks than a window has flags for
    assert lib.dbx_comm_create(0, 1, 0, 0, C.byref(h)) != 0             # empty window
    assert lib.dbx_comm_allreduce(None, None, 0, None, 0, None) != 0
    assert lib.dbx_comm_destroy(None) == 0


def test_sharded_upload_switches(monkeypatch):
    """distributed.sharded_upload_enabled: row-sharded upload of the input batch from 4 ranks on (measured: slower at 2), off with
    DBX_NO_SHARDED_UPLOAD=1, threshold overridable."""
    from deepbayes_b200 import distributed as dd
    monkeypatch.delenv("DBX_NO_SHARDED_UPLOAD", raising=False)
    monkeypatch.delenv("DBX_SHARDED_UPLOAD_MIN_RANKS", raising=False)
    assert [dd.sharded_upload_enabled(w) for w in (1, 2, 4, 8)] == [False, False, True, True]
    assert dd.sharded_upload_enabled() is False                       # no process group: one rank
    monkeypatch.setenv("DBX_SHARDED_UPLOAD_MIN_RANKS", "2")
    assert dd.sharded_upload_enabled(2)
    monkeypatch.setenv("DBX_NO_SHARDED_UPLOAD", "1")
    assert not dd.sharded_upload_enabled(8)
    # without a process group there is nothing to shard
    assert dd.sharded_upload(np.zeros((8, 3), np.float32), lambda a: a, "cpu") is None


def test_train_mode_and_attack_mode_arguments():
    """compile(..., train_mode=) is validated where it is used; the attacks' mode argument resolves like the forward path's."""
    import deepbayes_b200 as db
    from deepbayes_b200.analyzers import verifiers
    m = db.Sequential([db.Dense(8, activation="relu", input_shape=(16,)), db.Dense(4, activation="softmax")])
    opt = db.optimizers.BayesByBackprop().compile(m, None, batch_size=8)
    assert opt._train_mode() == "fp32"                                 # the reference's arithmetic unless asked otherwise
    opt = db.optimizers.BayesByBackprop().compile(m, None, batch_size=8, train_mode="bf16")
    assert opt._train_mode() == "bf16"
    opt.train_mode = "fp16"
    with pytest.raises(ValueError):
        opt._train_mode()

    class _M:
        mode = "fp32"
    assert verifiers._mode_of(_M()) == "fp32" and verifiers._mode_of(_M(), "bf16") == "bf16" and verifiers._mode_of(object()) == "bf16"


def test_massart_bound_check_replays_a_run_of_the_reference_loop(golden_dir):
    """tests/golden/reference_massart.npz = the reference's own massart_bound_check (verifiers.py:563-653, verify=True,
    classification=True) on the VOGN tutorial posterior, with the interval bounds every iteration saw.  The drop-in loop evaluates
    posterior samples in BLOCKS and walks the predicate / Clopper-Pearson / Massart halting rule over them on the host: fed the
    same per-sample bounds (a stand-in engine), it must stop after the same number of iterations with the same estimate -- all
    passes, all failures (the I[1] < 0.5 branch) and a mixed run.  okamoto_bound / absolute_massart_halting against the
    reference's values as well."""
    import torch
    from deepbayes_b200.analyzers import verifiers
    g = np.load(os.path.join(golden_dir, "reference_massart.npz"))
    for (e, d), want in zip(((0.1, 0.3), (0.05, 0.1), (0.25, 0.3)), g["okamoto"]):
        assert verifiers.okamoto_bound(e, d) == pytest.approx(want, rel=1e-12)
    for a, want in zip(g["halting_args"], g["halting"]):
        assert verifiers.absolute_massart_halting(int(a[0]), int(a[1]), [a[2], a[3]], a[4], a[5], a[6]) == int(want)

    class _Seq:
        def rows_and_outshape(self, a):
            return 1, (1, 10)

    class _Eng:
        C, K = 10, 784
        model = _Seq()

        def __init__(self, lo, hi):
            self.lo, self.hi, self.asked = lo, hi, []

        def ibp_samples(self, x, eps, nb, seed, s0, mode="fp32", **kw):
            self.asked.append((int(s0), int(nb), mode))
            idx = np.minimum(np.arange(s0, s0 + nb), len(self.lo) - 1)      # samples past the halting point are never looked at
            return torch.from_numpy(self.lo[idx].reshape(nb, 1, 10).astype(np.float32)), torch.from_numpy(self.hi[idx].reshape(nb, 1, 10).astype(np.float32))

    class _Model:
        seed, mode, det = 3, "bf16", False

        def __init__(self, eng):
            self._e, self._next = eng, 0

        def engine(self):
            return self._e

        def _take_ids(self, k):
            v = self._next
            self._next += k
            return v

        def predict(self, x, n=35):
            return np.zeros((1, 10), np.float32)

    for name in ("all_pass", "all_fail", "mixed"):
        eps, cls, conf, delta = g[name + "_params"]
        cls = int(cls)
        seen = []

        def predicate(worst, _cls=cls):
            assert np.asarray(worst).shape == (1, 10) and abs(float(np.sum(worst)) - 1.0) < 1e-5      # softmax of the worst-case logits
            ok = bool(np.argmax(np.squeeze(worst)) == _cls)
            seen.append(ok)
            return ok
        for block in (4096, 7):
            del seen[:]
            eng = _Eng(g[name + "_lo"], g[name + "_hi"])
            p, it, _ = verifiers.massart_bound_check(_Model(eng), g["x"], float(eps), predicate, cls=cls, classification=True, verify=True,
                                                     confidence=float(conf), delta=float(delta), block=block)
            want_p, want_it = g[name + "_result"]
            assert it == want_it and p == pytest.approx(want_p, rel=1e-12), (name, block, p, it)
            np.testing.assert_array_equal(np.asarray(seen, np.uint8), g[name + "_outcomes"])
            assert all(m == "fp32" for _, _, m in eng.asked)          # interval bounds default to fp32 whatever the model's forward mode


def test_train_learning_rate_decay_and_linear_epsilon_schedule():
    """optimizer.py:100-133: lrate_e = learning_rate / (1 + decay * e) and, with linear_schedule (the default, which also adds an
    epoch, :77-79), epsilon starts at 0 and grows by max_eps / epochs after every epoch (:105-108, :131-132).  The mini-batch
    loop of BayesByBackprop.train is driven with a recording `step` (no device work)."""
    import torch
    import deepbayes_b200 as db

    class Rec(db.optimizers.BayesByBackprop):
        def step(self, features, labels, lrate, noise=None, sync=True, eps_draws=None):
            self.rec.append((float(lrate), float(self.epsilon), len(features)))
            self._last_loss = [0.0, 0.0]
            self._last_probs = torch.zeros((len(features), 3))
            return self.posterior_mean, self.posterior_var

    m = db.Sequential([db.Dense(4, activation="relu", input_shape=(5,)), db.Dense(3, activation="softmax")])
    opt = Rec().compile(m, None, batch_size=4, learning_rate=0.2, decay=0.5, epochs=3, robust_train=2, epsilon=0.08)
    assert opt.epochs == 4 and opt.robust_linear
    opt.rec = []
    X = np.zeros((10, 5), np.float32)
    y = np.zeros(10, np.int64)
    opt.train(X, y, device_epochs=False)
    assert len(opt.rec) == 4 * 3 and [r[2] for r in opt.rec[:3]] == [4, 4, 2]                 # 3 mini-batches per epoch, ragged last one
    for e in range(4):
        for lr, eps, _ in opt.rec[3 * e:3 * e + 3]:
            assert lr == pytest.approx(0.2 / (1 + 0.5 * e)) and eps == pytest.approx(e * 0.08 / 4)
    assert opt.max_eps == pytest.approx(0.08) and opt.epsilon == pytest.approx(0.08)          # back at max_eps after the last epoch
    # without the schedule: epsilon constant, no extra epoch
    opt2 = Rec().compile(m, None, batch_size=4, learning_rate=0.2, epochs=2, robust_train=2, epsilon=0.08, linear_schedule=False)
    opt2.rec = []
    opt2.train(X, y, device_epochs=False)
    assert opt2.epochs == 2 and {r[1] for r in opt2.rec} == {0.08} and {r[0] for r in opt2.rec} == {0.2}


def test_stored_sample_draws_consume_the_numpy_stream_like_the_reference_loop():
    """posteriormodel.py:77 draws ONE index per loop iteration (`np.random.choice(range(S), p=frequency)`); the drop-in
    PosteriorModel draws the n indices of a predict in one call (`np.random.choice(S, size=n, p=frequency)`).  Under the same
    global seed both give the same index sequence and leave the stream in the same state."""
    S, n = 17, 40
    f = np.random.Generator(np.random.Philox(1)).random(S)
    f = f / f.sum()
    np.random.seed(123)
    one_by_one = [int(np.random.choice(range(S), p=f)) for _ in range(n)]
    after_a = np.random.random()
    np.random.seed(123)
    at_once = np.random.choice(S, size=n, p=f).astype(np.int32)
    after_b = np.random.random()
    assert one_by_one == at_once.tolist() and after_a == after_b
