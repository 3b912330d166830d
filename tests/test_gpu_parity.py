"""GPU parity tests: the CUDA path, called through the C ABI (deepbayes_b200.engine.Engine is
a thin ctypes shim), against the CPU oracle on identical seeded inputs with the oracle's eps
tensors injected, and against the committed golden outputs.

Tolerances (BASELINE.json north_star): fp32 mode <= 1e-5 relative; bf16 mode <= 2e-2 absolute
on predictive probabilities; argmax of the predictive mean identical."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import cases  # noqa: E402

pytestmark = pytest.mark.gpu

FP32_RTOL, FP32_ATOL = 1e-5, 2e-7
BF16_ATOL = 2e-2


def _model_for(name, c):
    import deepbayes_b200 as db
    L = c["layers"]
    layers = []
    first = True
    for l in L:
        kw = {"input_shape": c["in_shape"]} if first else {}
        if l.kind == "dense":
            layers.append(db.Dense(l.units, activation=l.activation, **kw))
        elif l.kind == "conv2d":
            layers.append(db.Conv2D(l.cout, (l.kh, l.kw), l.strides, l.padding, l.activation, **kw))
        elif l.kind == "maxpool2d":
            layers.append(db.MaxPooling2D(l.pool, l.strides))
        else:
            layers.append(db.Flatten())
        first = False
    return db.Sequential(layers)


@pytest.fixture(scope="module")
def golden(golden_dir):
    return np.load(os.path.join(golden_dir, "oracle_cases.npz"))


@pytest.fixture(scope="module")
def ctx():
    cache = {}

    def get(name):
        if name not in cache:
            from deepbayes_b200.engine import Engine
            c = cases.case_inputs(name)
            m = _model_for(name, c)
            eng = Engine(m)
            eng.set_gaussian(c["mean"], c["std"])
            from oracle import oracle as o
            c["eps_flat"] = np.stack([o.flatten_list(e) for e in c["eps"]])
            cache[name] = (c, m, eng)
        return cache[name]
    return get


def test_native_library_is_loaded():
    from deepbayes_b200 import _lib
    _lib.load()
    assert any("libdeepbayes_b200.so" in l for l in open("/proc/self/maps").read().splitlines())


def test_philox_bit_exact_and_normals(oracle):
    import deepbayes_b200 as db
    from deepbayes_b200.engine import Engine
    m = db.Sequential([db.Dense(7, activation="relu", input_shape=(13,)), db.Dense(3, activation="softmax")])
    eng = Engine(m)
    for seed, s in ((0, 0), (12345678901234, 7), (2 ** 63 + 5, 2 ** 33 + 1)):
        got = eng.philox_raw(seed, s, 1001)
        np.testing.assert_array_equal(got, oracle.philox_raw(seed, s, 1001))
    P = eng.P
    eng.set_gaussian([np.zeros(s, np.float32) for s in m.weight_shapes()], [np.ones(s, np.float32) for s in m.weight_shapes()])
    W = eng.sample(3, seed=99, s0=5).cpu().numpy()  # mu=0, sigma=1 -> W is the eps stream itself
    for i in range(3):
        np.testing.assert_allclose(W[i], oracle.philox_normals(99, 5 + i, P), rtol=0, atol=5e-6)
    # same global sample id -> same draw, independent of the call split
    np.testing.assert_array_equal(eng.sample(1, seed=99, s0=6).cpu().numpy()[0], W[1])


def test_philox_statistics(oracle):
    import deepbayes_b200 as db
    from deepbayes_b200.engine import Engine
    m = db.Sequential([db.Dense(512, activation="relu", input_shape=(784,)), db.Dense(10, activation="softmax")])
    eng = Engine(m)
    eng.set_gaussian([np.zeros(s, np.float32) for s in m.weight_shapes()], [np.ones(s, np.float32) for s in m.weight_shapes()])
    z = eng.sample(8, seed=3, s0=0).cpu().numpy().astype(np.float64).reshape(-1)
    assert abs(z.mean()) < 2e-3 and abs(z.std() - 1) < 2e-3
    assert abs((z ** 3).mean()) < 1e-2 and abs((z ** 4).mean() - 3) < 2e-2
    assert np.abs(np.corrcoef(z[:-1], z[1:])[0, 1]) < 2e-3


@pytest.mark.parametrize("name", cases.NAMES)
def test_sample_matches_oracle(ctx, golden, oracle, name):
    c, m, eng = ctx(name)
    W = eng.sample(c["n"], eps=c["eps_flat"]).cpu().numpy()
    want = np.stack([oracle.flatten_list(oracle.sample_gaussian(c["mean"], c["std"], e)) for e in c["eps"]])
    # fmaf(sigma, eps, mu) vs the reference's float64 form then fp32 cast: <= 1 ulp (double rounding)
    np.testing.assert_allclose(W, want, rtol=1.2e-7, atol=1e-12)
    assert (W == want).mean() > 0.999
    np.testing.assert_allclose(W[:, ::cases.w_stride(c["layers"])], golden[name + "/sample_w"], rtol=1.2e-7, atol=1e-12)


@pytest.mark.parametrize("name", cases.NAMES)
def test_predict_fp32(ctx, golden, oracle, name):
    c, m, eng = ctx(name)
    n = c["n"]
    s1, s2 = eng.predict_sums(c["x"], n, eps=c["eps_flat"], mode="fp32", second_moment=True)
    got = (s1 / float(n)).cpu().numpy()
    want = oracle.predict(c["layers"], c["mean"], c["std"], c["x"], n, c["eps"])
    np.testing.assert_allclose(got, want, rtol=FP32_RTOL, atol=FP32_ATOL)
    np.testing.assert_allclose(got, golden[name + "/predict_fp32"], rtol=FP32_RTOL, atol=FP32_ATOL)
    np.testing.assert_allclose(s1.cpu().numpy(), golden[name + "/sum1_fp32"], rtol=FP32_RTOL, atol=n * FP32_ATOL)
    np.testing.assert_allclose(s2.cpu().numpy(), golden[name + "/sum2_fp32"], rtol=2 * FP32_RTOL, atol=n * FP32_ATOL)
    assert eng.last_path() == ("stream" if (c["B"] <= 16 and name != "cnn_small") else "generic")
    lg, _ = eng.predict_sums(c["x"], n, eps=c["eps_flat"], mode="fp32", out_kind="logits")
    np.testing.assert_allclose((lg / float(n)).cpu().numpy(), golden[name + "/predict_logits_fp32"], rtol=FP32_RTOL, atol=2e-6)
    inf, _ = eng.predict_sums(c["x"], n, eps=c["eps_flat"], mode="fp32", inflate=2.0)
    np.testing.assert_allclose((inf / float(n)).cpu().numpy(), golden[name + "/predict_inflate2_fp32"], rtol=FP32_RTOL, atol=FP32_ATOL)


@pytest.mark.parametrize("name", cases.NAMES)
@pytest.mark.parametrize("fused", [True, False])
def test_predict_bf16(ctx, golden, oracle, name, fused, dbx_opt):
    c, m, eng = ctx(name)
    dbx_opt("NO_FUSED", "0" if fused else "1")
    n = c["n"]
    s1, _ = eng.predict_sums(c["x"], n, eps=c["eps_flat"], mode="bf16")
    got = (s1 / float(n)).cpu().numpy()
    fusable = name in ("mlp_cfg1", "mlp_cfg2")
    small = c["B"] <= 16 and name != "cnn_small"      # small batches take the fp32 streaming kernel in any mode
    want = ("stream",) if small else ("fused_tcgen05",) if (fused and fusable) else ("generic", "generic_tc")
    assert eng.last_path() in want
    want_bf = golden[name + "/predict_bf16"]   # oracle with bf16-rounded GEMM operands
    want_32 = golden[name + "/predict_fp32"]
    assert np.abs(got - want_32).max() <= BF16_ATOL
    # against the bf16-emulating oracle only accumulation order differs
    np.testing.assert_allclose(got, want_bf, rtol=0, atol=2e-3)
    if c["layers"][-1].activation == "softmax":
        margin = np.sort(want_32, axis=-1)
        clear = (margin[:, -1] - margin[:, -2]) > 2 * BF16_ATOL
        assert (got.argmax(-1) == want_32.argmax(-1))[clear].all()
        assert (got.argmax(-1) == want_bf.argmax(-1)).all(), "argmax differs from the bf16 oracle"


@pytest.mark.parametrize("name", ["mlp_small", "mlp_cfg2", "cnn_small"])
def test_stored_posterior(ctx, golden, oracle, name):
    c, m, eng = ctx(name)
    n = c["n"]
    samples = [oracle.sample_gaussian(c["mean"], c["std"], e) for e in c["eps"]]
    slab = np.stack([oracle.flatten_list(s) for s in samples])
    eng.set_samples(slab)
    freq = oracle.normalise_freq(np.arange(1, n + 1)).astype(np.float32)
    s1, _ = eng.predict_stored_sums(c["x"], n, j0=0, wts=freq, mode="fp32")
    np.testing.assert_allclose(s1.cpu().numpy(), golden[name + "/stored_expect_fp32"], rtol=FP32_RTOL, atol=FP32_ATOL)
    idx = [(3 * j + 1) % n for j in range(2 * n)]
    s1, _ = eng.predict_stored_sums(c["x"], len(idx), idx=idx, mode="fp32")
    np.testing.assert_allclose((s1 / float(len(idx))).cpu().numpy(), golden[name + "/stored_mc_fp32"], rtol=FP32_RTOL, atol=FP32_ATOL)
    s1b, _ = eng.predict_stored_sums(c["x"], len(idx), idx=idx, mode="bf16")
    assert np.abs((s1b / float(len(idx))).cpu().numpy() - golden[name + "/stored_mc_fp32"]).max() <= BF16_ATOL
    # range check: asking past the slab fails loudly
    from deepbayes_b200._lib import DbxError
    with pytest.raises((DbxError, ValueError)):
        eng.predict_stored_sums(c["x"], n + 1, j0=0, mode="fp32")


@pytest.mark.parametrize("name", ["mlp_small", "mlp_cfg1", "cnn_small"])
def test_bbb_forward(ctx, golden, oracle, name):
    c, m, eng = ctx(name)
    rho = [oracle.inverse_softplus(s) for s in c["std"]]
    eng.set_gaussian(c["mean"], rho, rho=True)
    try:
        out, W, noise = eng.bbb_forward(c["x"], eps=c["eps_flat"][0], mode="fp32", return_weights=True)
        want, w = oracle.bbb_forward(c["layers"], c["mean"], rho, c["eps"][0], c["x"])
        # W = mean + softplus(rho)*eps formed in fp32 in the reference's order: device softplus vs numpy <= 2 ulp
        np.testing.assert_allclose(W.cpu().numpy(), oracle.flatten_list(w), rtol=3e-7, atol=6e-8)
        np.testing.assert_array_equal(noise.cpu().numpy(), c["eps_flat"][0])
        np.testing.assert_allclose(out.cpu().numpy(), want, rtol=FP32_RTOL, atol=FP32_ATOL)
        np.testing.assert_allclose(out.cpu().numpy(), golden[name + "/bbb_fp32"], rtol=FP32_RTOL, atol=FP32_ATOL)
        out2 = eng.bbb_forward(c["x"], eps=c["eps_flat"][0], mode="fp32")
        np.testing.assert_allclose(out2.cpu().numpy(), out.cpu().numpy(), rtol=1e-6, atol=1e-7)
        mu, sg = eng.get_gaussian()
        np.testing.assert_allclose(sg, oracle.flatten_list([oracle.softplus(r) for r in rho]), rtol=3e-7)
    finally:
        eng.set_gaussian(c["mean"], c["std"])


@pytest.mark.parametrize("name", ["mlp_small", "mlp_cfg2", "cnn_small"])
@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_count_predicate(ctx, golden, oracle, name, mode):
    c, m, eng = ctx(name)
    counts, flags = eng.count_predicate(c["x"], c["labels"], c["n"], eps=c["eps_flat"], mode=mode, return_flags=True)
    flags, counts = flags.cpu().numpy(), counts.cpu().numpy()
    np.testing.assert_array_equal(counts, flags.sum(0))
    if mode == "fp32":
        want_flags = golden[name + "/flags"]
    else:
        want_flags, _ = oracle.count_predicate(c["layers"], c["mean"], c["std"], c["x"], c["labels"], c["n"], c["eps"], bf16=True)
    # bit-exact except where the top-2 logits are within rounding of each other
    assert (flags != want_flags).mean() <= (0.0 if mode == "fp32" else 0.02)


def test_edge_cases(oracle):
    import deepbayes_b200 as db
    from deepbayes_b200.engine import Engine
    from deepbayes_b200._lib import DbxError
    m = db.Sequential([db.Dense(3, activation="softmax", input_shape=(5,))])
    eng = Engine(m)
    with pytest.raises(DbxError):  # posterior not set
        eng.predict_sums(np.zeros((1, 5), np.float32), 1, mode="fp32")
    L = oracle.mlp([5, 3])
    mean, std = oracle.synthetic_posterior(L, seed=1)
    eng.set_gaussian(mean, std)
    with pytest.raises(ValueError):
        eng.set_gaussian(mean[:1], std[:1])
    # B = 1, n = 1, single layer
    x = oracle.synthetic_x((1, 5), seed=2)
    eps = oracle.synthetic_eps(L, 1, seed=3)
    s1, _ = eng.predict_sums(x, 1, eps=oracle.flatten_list(eps[0])[None], mode="fp32")
    np.testing.assert_allclose(s1.cpu().numpy(), oracle.predict(L, mean, std, x, 1, eps), rtol=FP32_RTOL, atol=FP32_ATOL)
    # sigma = 0 -> every sample equals the deterministic forward, for any seed
    eng.set_gaussian(mean, [np.zeros_like(s) for s in std])
    s1, s2 = eng.predict_sums(x, 9, seed=77, mode="fp32", second_moment=True)
    det = oracle.forward(L, mean, x)
    np.testing.assert_allclose((s1 / 9.0).cpu().numpy(), det, rtol=FP32_RTOL, atol=FP32_ATOL)
    np.testing.assert_allclose((s2 / 9.0).cpu().numpy(), det ** 2, rtol=3e-5, atol=FP32_ATOL)
    with pytest.raises(DbxError):
        eng.predict_sums(x, 0, mode="fp32")


def test_sample_split_independence(ctx):
    """Global sample ids: n samples in one call == the same ids split over two calls (this is
    what makes the multi-GPU shard sum independent of G)."""
    c, m, eng = ctx("mlp_cfg1")
    for mode, tol in (("fp32", 1e-6), ("bf16", 1e-6)):
        a, _ = eng.predict_sums(c["x"], 10, seed=5, s0=100, mode=mode)
        b1, _ = eng.predict_sums(c["x"], 4, seed=5, s0=100, mode=mode)
        b2, _ = eng.predict_sums(c["x"], 6, seed=5, s0=104, mode=mode)
        np.testing.assert_allclose(a.cpu().numpy(), (b1 + b2).cpu().numpy(), rtol=1e-5, atol=tol)
    # the interval forward shards the same way (chernoff_bound_verification with model.shard)
    for mode in ("fp32", "bf16"):
        a = eng.ibp_predict(c["x"], 0.01, c["labels"], 10, seed=5, s0=100, mode=mode, want=("worst", "counts"))
        b1 = eng.ibp_predict(c["x"], 0.01, c["labels"], 4, seed=5, s0=100, mode=mode, want=("worst", "counts"))
        b2 = eng.ibp_predict(c["x"], 0.01, c["labels"], 6, seed=5, s0=104, mode=mode, want=("worst", "counts"))
        np.testing.assert_allclose(a["worst"].cpu().numpy(), (b1["worst"] + b2["worst"]).cpu().numpy(), rtol=1e-5, atol=1e-6)
        np.testing.assert_array_equal(a["counts"].cpu().numpy(), (b1["counts"] + b2["counts"]).cpu().numpy())
    cnt = eng.count_predicate(c["x"], c["labels"], 64, seed=5, s0=0, mode="bf16").cpu().numpy()
    c1 = eng.count_predicate(c["x"], c["labels"], 24, seed=5, s0=0, mode="bf16").cpu().numpy()
    c2 = eng.count_predicate(c["x"], c["labels"], 40, seed=5, s0=24, mode="bf16").cpu().numpy()
    np.testing.assert_array_equal(cnt, c1 + c2)


@pytest.mark.parametrize("cm", ["single", "pair", "pair_qs"])
@pytest.mark.parametrize("dims,B,n", [([784, 512, 512, 10], 1000, 70), ([300, 512, 512, 16], 5000, 9), ([784, 512, 10], 700, 33), ([784, 128, 10], 300, 20),
                                       ([100, 256, 256, 7], 513, 40), ([784, 384, 384, 10], 260, 12)])
def test_fused_vs_generic_device_crosscheck(oracle, dbx_opt, cm, dims, B, n):
    """The tcgen05 fused kernels (single-CTA and CTA-pair, both work orders, ragged row tiles, several chunks
    and work segments) against the plain CUDA-core bf16 path on the same Philox draws: both round
    the GEMM operands to bf16 and accumulate in fp32, so they agree to accumulation order."""
    import deepbayes_b200 as db
    from deepbayes_b200.engine import Engine
    L = oracle.mlp(dims)
    layers = [db.Dense(dims[1], activation="relu", input_shape=(dims[0],))]
    for h in dims[2:-1]:
        layers.append(db.Dense(h, activation="relu"))
    layers.append(db.Dense(dims[-1], activation="softmax"))
    eng = Engine(db.Sequential(layers))
    mean, std = oracle.synthetic_posterior(L, seed=5, sigma_scale=0.5)
    eng.set_gaussian(mean, std)
    x = oracle.synthetic_x((B, dims[0]), seed=6)
    dbx_opt("NO_FUSED", "1")
    dbx_opt("NO_TC", "1")     # reference = plain CUDA-core bf16 kernels
    g1, g2 = eng.predict_sums(x, n, seed=17, s0=3, mode="bf16", second_moment=True)
    assert eng.last_path() == "generic"
    dbx_opt("NO_TC", "0")
    t1, t2 = eng.predict_sums(x, n, seed=17, s0=3, mode="bf16", second_moment=True)   # per-layer tcgen05 GEMMs
    assert eng.last_path() == "generic_tc"   # at least the hidden layers are GEMM-eligible (widths % 8 == 0)
    np.testing.assert_allclose((t1 / n).cpu().numpy(), (g1 / n).cpu().numpy(), rtol=0, atol=1.5e-3)
    dbx_opt("NO_FUSED", "0")
    # pair = fused_mlp2_kernel (cta_group::2, the default for B > 128), work order [s][q]; pair_qs = the same kernel with the
    # [q][s] contiguous-range order; single = fused_mlp_kernel<1> (the kernel a lone 128-row tile gets) forced on every shape
    dbx_opt("FUSED_KERNEL", 1 if cm == "single" else 2)
    dbx_opt("FUSED_ORDER", 0 if cm.endswith("_qs") else 1)
    dbx_opt("FUSED_NS", "16")   # several chunks -> accumulate path of finish_kernel
    for rep in range(3):
        f1, f2 = eng.predict_sums(x, n, seed=17, s0=3, mode="bf16", second_moment=True)
        assert eng.last_path() == "fused_tcgen05"
        np.testing.assert_allclose((f1 / n).cpu().numpy(), (g1 / n).cpu().numpy(), rtol=0, atol=1.5e-3)
        np.testing.assert_allclose((f2 / n).cpu().numpy(), (g2 / n).cpu().numpy(), rtol=0, atol=1.5e-3)
        if rep == 0:
            first = f1.clone()
        else:  # deterministic: bit-identical run to run
            assert bool((f1 == first).all())
    lab = np.random.Generator(np.random.Philox(1)).integers(0, dims[-1], size=B).astype(np.int32)
    cf, ff = eng.count_predicate(x, lab, n, seed=17, s0=3, mode="bf16", return_flags=True)
    dbx_opt("NO_FUSED", "1")
    cg, fg = eng.count_predicate(x, lab, n, seed=17, s0=3, mode="bf16", return_flags=True)
    assert (ff != fg).float().mean().item() < 5e-3
    np.testing.assert_array_equal(cf.cpu().numpy(), ff.sum(0).cpu().numpy())


def test_fused_long_run_chunking(oracle, dbx_opt):
    """Small batches with many samples run the fused kernel in launches of up to 1024 samples (instead of 256): same
    predicate counts, same sums up to the accumulation order."""
    L, eng, mean, std = _mlp_engine([784, 512, 512, 10], oracle, sigma_scale=0.3)
    B, n = 100, 2304 + 37
    x = oracle.synthetic_x((B, 784), seed=12)
    labels = np.random.Generator(np.random.Philox(8)).integers(0, 10, size=B).astype(np.int32)
    dbx_opt("FUSED_NS", "256")
    c0 = eng.count_predicate(x, labels, n, seed=5, mode="bf16").cpu().numpy()
    s0, q0 = eng.predict_sums(x, n, seed=5, mode="bf16", second_moment=True)
    dbx_opt("FUSED_NS", None)
    c1 = eng.count_predicate(x, labels, n, seed=5, mode="bf16").cpu().numpy()
    s1, q1 = eng.predict_sums(x, n, seed=5, mode="bf16", second_moment=True)
    assert eng.last_path() == "fused_tcgen05"
    np.testing.assert_array_equal(c0, c1)
    np.testing.assert_allclose(s0.cpu().numpy() / n, s1.cpu().numpy() / n, rtol=0, atol=2e-6)
    np.testing.assert_allclose(q0.cpu().numpy() / n, q1.cpu().numpy() / n, rtol=0, atol=2e-6)


@pytest.mark.parametrize("dims,B", [([784, 1024, 1024, 10], 16), ([784, 1024, 1024, 10], 1), ([784, 512, 10], 5),
                                    ([37, 50, 10], 7), ([13, 24, 16, 3], 9), ([64, 40, 12], 2)])
def test_stream_vs_generic_device_crosscheck(oracle, dbx_opt, dims, B):
    """Small-batch streaming kernel (weights read / generated once, fp32) against the generic fp32
    kernels: Philox draws, injected eps, stored samples with index list and weights, count sink."""
    import deepbayes_b200 as db
    from deepbayes_b200.engine import Engine
    L = oracle.mlp(dims)
    layers = [db.Dense(dims[1], activation="relu", input_shape=(dims[0],))]
    for h in dims[2:-1]:
        layers.append(db.Dense(h, activation="relu"))
    layers.append(db.Dense(dims[-1], activation="softmax"))
    eng = Engine(db.Sequential(layers))
    mean, std = oracle.synthetic_posterior(L, seed=5, sigma_scale=0.5)
    eng.set_gaussian(mean, std)
    x = oracle.synthetic_x((B, dims[0]), seed=6)
    n = 23
    eps = np.random.Generator(np.random.Philox(4)).standard_normal((n, eng.P)).astype(np.float32)
    slab = eng.sample(9, seed=3, s0=0)
    eng.set_samples(slab)
    idx = [(5 * j + 2) % 9 for j in range(n)]
    wts = np.linspace(0.1, 1.0, 9).astype(np.float32)
    lab = np.random.Generator(np.random.Philox(1)).integers(0, dims[-1], size=B).astype(np.int32)
    res = {}
    for tag, env in (("generic", "1"), ("stream", "0")):
        dbx_opt("NO_STREAM", env)
        dbx_opt("NO_FUSED", "1")
        r = {}
        r["philox"] = eng.predict_sums(x, n, seed=17, s0=3, mode="fp32", second_moment=True)
        assert eng.last_path() == tag
        r["eps"] = eng.predict_sums(x, n, eps=eps, mode="fp32", out_kind="logits", second_moment=True)
        r["stored_idx"] = eng.predict_stored_sums(x, n, idx=idx, mode="fp32", second_moment=True)
        r["stored_wts"] = eng.predict_stored_sums(x, 9, j0=0, wts=wts, mode="fp32", second_moment=True)
        r["count"] = eng.count_predicate(x, lab, n, seed=17, s0=3, mode="fp32", return_flags=True)
        assert eng.last_path() == tag
        res[tag] = r
    for k in ("philox", "eps", "stored_idx", "stored_wts"):
        for a, b in zip(res["stream"][k], res["generic"][k]):
            np.testing.assert_allclose(a.cpu().numpy(), b.cpu().numpy(), rtol=2e-5, atol=2e-5 if k == "eps" else 2e-6)
    assert (res["stream"]["count"][1] != res["generic"]["count"][1]).float().mean().item() < 0.02
    np.testing.assert_array_equal(res["stream"]["count"][0].cpu().numpy(), res["stream"]["count"][1].sum(0).cpu().numpy())
    # bf16 mode at small B is served by the same fp32 streaming kernel
    dbx_opt("NO_STREAM", "0")
    b1, _ = eng.predict_sums(x, n, seed=17, s0=3, mode="bf16")
    assert eng.last_path() == "stream"
    np.testing.assert_array_equal(b1.cpu().numpy(), res["stream"]["philox"][0].cpu().numpy())


def test_mc_converges_to_oracle_mc(ctx, oracle):
    """Philox path (no injected eps): the MC mean over many samples agrees with the oracle's
    MC mean over its own RNG to within the Monte-Carlo standard error."""
    c, m, eng = ctx("mlp_small")
    n = 4096
    s1, s2 = eng.predict_sums(c["x"], n, seed=11, mode="fp32", second_moment=True)
    mean = (s1 / n).cpu().numpy()
    var = np.maximum((s2 / n).cpu().numpy() - mean ** 2, 0)
    eps = oracle.synthetic_eps(c["layers"], 512, seed=99)
    ref = oracle.predict(c["layers"], c["mean"], c["std"], c["x"], 512, eps)
    se = np.sqrt(var / n + var / 512) + 1e-4
    assert (np.abs(mean - ref) < 6 * se).all()


def test_posterior_model_api(golden_dir, oracle):
    """Drop-in surface on the reference's own saved posterior (loaded from its TF pickles)."""
    import deepbayes_b200 as db
    p = os.path.join(golden_dir, "posteriors", "VOGN_MNIST_Posterior")
    L = oracle.mlp([784, 128, 10])
    x = oracle.synthetic_x((6, 1, 784), seed=0).astype(np.float64)  # tutorials pass float64 (B,1,784)
    np.random.seed(0)
    pm = db.PosteriorModel(p, mode="fp32")
    assert pm.det is False and pm.sample_based is False and pm.input_upper == np.inf
    out = pm.predict(x, n=35)
    assert out.shape == (6, 1, 10) and out.dtype == np.float32
    det = oracle.forward(L, pm.posterior_mean, x.reshape(6, 784)).reshape(6, 1, 10)
    assert np.abs(out - det).max() < 2e-3 and np.allclose(out.sum(-1), 1, rtol=1e-5)
    # predict leaves the model holding the last sample (posteriormodel.py:96); sample() does not touch it
    w_after = pm.model.get_weights()
    assert not np.array_equal(w_after[0], pm.posterior_mean[0])
    assert np.abs(w_after[0] - pm.posterior_mean[0]).max() < 6 * pm.posterior_var[0].max()
    s = pm.sample()
    assert [a.shape for a in s] == [(784, 128), (128,), (128, 10), (10,)]
    assert all(a.dtype == np.float64 for a in s)          # np.random.normal at posteriormodel.py:74-75 returns float64
    np.testing.assert_array_equal(pm.model.get_weights()[0], w_after[0])
    np.testing.assert_allclose(pm._predict(x), oracle.forward(L, w_after, x.reshape(6, 784)).reshape(6, 1, 10), rtol=FP32_RTOL, atol=FP32_ATOL)
    lg = pm.predict_logits(x, n=5)
    assert lg.shape == (6, 1, 10) and np.abs(lg - oracle.forward(L, pm.posterior_mean, x.reshape(6, 784), out="logits").reshape(6, 1, 10)).max() < 0.05
    wl = pm.model.get_weights()
    nb = pm._predict_logits(x)  # reference bug kept: no bias
    np.testing.assert_allclose(nb, oracle.forward(L, wl[:-1] + [np.zeros_like(wl[-1])], x.reshape(6, 784), out="logits").reshape(6, 1, 10), rtol=1e-5, atol=2e-6)
    # deterministic mode short-circuits (posteriormodel.py:42-43,92-93)
    pd = db.PosteriorModel(p, deterministic=True, mode="fp32")
    np.testing.assert_allclose(pd.predict(x, n=50), det, rtol=FP32_RTOL, atol=FP32_ATOL)
    np.testing.assert_array_equal(pd.sample()[0], pd.posterior_mean[0])
    # bf16 default mode agrees within the stated tolerance, argmax identical
    pb = db.PosteriorModel(p, seed=pm.seed)
    ob = pb.predict(x, n=35)
    assert np.abs(ob - out).max() < BF16_ATOL and (ob.argmax(-1) == out.argmax(-1)).all()
    # analyzers on top
    ep, al = db.analyzers.variational_uncertainty(pm, x, num_samples=64)
    assert ep.shape == (6, 1, 10) and (ep > -1e-6).all() and (al > -1e-6).all()
    assert len(db.analyzers.predictive_entropy(pm, x, num_models=8)) == 6


def test_sample_based_posterior_model(tmp_path, oracle):
    import deepbayes_b200 as db
    L = oracle.mlp([20, 16, 4])
    mean, std = oracle.synthetic_posterior(L, seed=1, sigma_scale=0.5)
    m = db.Sequential([db.Dense(16, activation="relu", input_shape=(20,)), db.Dense(4, activation="softmax")])
    hmc = db.optimizers.HamiltonianMonteCarlo().compile(m, None)
    hmc.posterior_mean = mean
    samples = [oracle.sample_gaussian(mean, std, e) for e in oracle.synthetic_eps(L, 6, seed=2)]
    for i, s in enumerate(samples):
        hmc.record(s, multiplicity=i + 1)
    hmc.save(str(tmp_path / "hmc"))
    pm = db.PosteriorModel(str(tmp_path / "hmc"), mode="fp32")
    assert pm.sample_based and pm.num_post_samps == 6 and abs(pm.frequency.sum() - 1) < 1e-12
    x = oracle.synthetic_x((5, 20), seed=3)
    np.random.seed(42)
    got = pm.predict(x, n=40)
    np.random.seed(42)
    idx = np.random.choice(6, size=40, p=pm.frequency)
    want = oracle.predict_stored(L, samples, x, list(idx)) / np.float32(40.0)
    np.testing.assert_allclose(got, want, rtol=FP32_RTOL, atol=FP32_ATOL)
    np.testing.assert_array_equal(pm.model.get_weights()[0], samples[idx[-1]][0])


def test_bbb_optimizer_api(oracle):
    import deepbayes_b200 as db
    L = oracle.mlp([30, 24, 10])
    m = db.Sequential([db.Dense(24, activation="relu", input_shape=(30,)), db.Dense(10, activation="softmax")])
    opt = db.optimizers.BayesByBackprop().compile(m, None, batch_size=16, seed=4)
    x = oracle.synthetic_x((16, 30), seed=0)
    eps = oracle.synthetic_eps(L, 1, seed=1)[0]
    pred, ws, noise = opt.forward_sample(x, eps=eps, return_weights=True)
    want, w = oracle.bbb_forward(L, opt.posterior_mean, opt.posterior_var, eps, x)
    np.testing.assert_allclose(pred, want, rtol=FP32_RTOL, atol=FP32_ATOL)
    for a, b in zip(ws, w):
        np.testing.assert_allclose(a, b, rtol=3e-7, atol=6e-8)
    np.testing.assert_array_equal(m.get_weights()[0], ws[0])         # step leaves the sample in the model (:59)
    np.testing.assert_allclose(opt.predict(x), want, rtol=FP32_RTOL, atol=FP32_ATOL)  # single forward, no sampling (:293-298)
    s = opt.sample()
    assert all(a.dtype == np.float64 for a in s)          # optimizer.py:199-201 / bayesbybackprop.py:178-181
    sd = np.sqrt(1 / 30.0)
    assert abs(np.std(s[0]) - sd) < 0.1 * sd                          # N(0, softplus(rho)) = N(0, prior std)
    p2 = opt.forward_sample(x)                                        # Philox noise
    assert p2.shape == (16, 10) and np.allclose(p2.sum(-1), 1, rtol=1e-5)


def test_statistical_verification(ctx, oracle):
    import deepbayes_b200 as db
    p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "posteriors", "VOGN_MNIST_Posterior")
    pm = db.PosteriorModel(p, mode="fp32", seed=3)
    x = oracle.synthetic_x((4, 784), seed=0)
    cls = pm.predict(x, n=16).argmax(-1)
    res = db.analyzers.statistical_robustness(pm, x, cls, return_flags=True)
    assert res["n"] == 380 and res["flags"].shape == (380, 4)
    assert (res["p"] > 0.9).all()
    frac, iters, _ = db.analyzers.massart_bound_check(pm, x[:1], 0.0, cls=int(cls[0]), adv=x[:1], verify=False)
    flags = np.ones(1000, np.uint8)
    want_frac, want_iters = oracle.massart_loop(flags)
    if frac == 1.0:
        assert iters == want_iters


def test_predict_host_e2e(ctx, oracle):
    c, m, eng = ctx("mlp_cfg1")
    mean = eng.predict_host(c["x"], 8, seed=21, s0=0, mode="fp32")
    s1, _ = eng.predict_sums(c["x"], 8, seed=21, s0=0, mode="fp32")
    np.testing.assert_allclose(mean, (s1 / 8.0).cpu().numpy(), rtol=1e-6, atol=1e-7)
    mean2, var2 = eng.predict_host(c["x"], 8, seed=21, s0=0, mode="bf16", variance=True)
    assert np.abs(mean2 - mean).max() < BF16_ATOL and (var2 > -1e-6).all()


def test_reference_run_vectors(golden_dir, oracle):
    """CUDA path against the outputs of the reference's own source (tests/golden/reference_run.npz),
    with the reference's global-RNG draws injected as eps."""
    import deepbayes_b200 as db
    ref = np.load(os.path.join(golden_dir, "reference_run.npz"))
    p = os.path.join(golden_dir, "posteriors", "VOGN_MNIST_Posterior")
    pm = db.PosteriorModel(p, mode="fp32")
    x = oracle.synthetic_x((6, 1, 784), seed=0).astype(np.float64)
    np.random.seed(20240601)
    eps = np.stack([np.concatenate([np.random.standard_normal(t.shape).reshape(-1) for t in pm.posterior_mean]) for _ in range(7)])
    s1, _ = pm._engine.predict_sums(x, 7, eps=eps.astype(np.float32), mode="fp32")
    np.testing.assert_allclose((s1 / 7.0).cpu().numpy().reshape(6, 1, 10), ref["pm_predict_n7"], rtol=FP32_RTOL, atol=FP32_ATOL)
    sl, _ = pm._engine.predict_sums(x, 5, eps=eps[:5].astype(np.float32), mode="fp32", out_kind="logits")
    np.testing.assert_allclose((sl / 5.0).cpu().numpy().reshape(6, 1, 10), ref["pm_predict_logits_n5"], rtol=FP32_RTOL, atol=3e-6)
    W = pm._engine.sample(1, eps=eps[6:7].astype(np.float32)).cpu().numpy()[0]
    o1 = 784 * 128 + 128
    np.testing.assert_allclose(W[o1:o1 + 1280].reshape(128, 10), ref["pm_weights_after_predict_W1"], rtol=1.2e-7, atol=1e-12)
    pd = db.PosteriorModel(p, deterministic=True, mode="fp32")
    np.testing.assert_allclose(pd.predict(x, n=50), ref["pm_det_predict"], rtol=FP32_RTOL, atol=FP32_ATOL)
    pm.model.set_weights(pm.posterior_mean)
    np.testing.assert_allclose(pm._predict_logits(x), ref["pm__predict_logits_nobias"], rtol=FP32_RTOL, atol=3e-6)
    # Bayes-by-Backprop: compile + the sampling/forward half of step with the reference's noise
    m = db.Sequential([db.Dense(24, activation="relu", input_shape=(30,)), db.Dense(10, activation="softmax")])
    opt = db.optimizers.BayesByBackprop().compile(m, None, batch_size=16, inflate_prior=2.0)
    assert opt.epochs == int(ref["bbb_epochs_after_compile"][0])
    for i in range(4):
        np.testing.assert_array_equal(opt.prior_var[i], ref["bbb_prior_var_%d" % i])
        np.testing.assert_allclose(opt.posterior_var[i], ref["bbb_rho_%d" % i], rtol=1e-6)
    noise = [ref["bbb_step_noise_%d" % i] for i in range(4)]
    pred, ws, _ = opt.forward_sample(oracle.synthetic_x((16, 30), seed=3), eps=noise, return_weights=True)
    np.testing.assert_allclose(pred, ref["bbb_step_predictions"], rtol=FP32_RTOL, atol=FP32_ATOL)
    np.testing.assert_allclose(ws[0], ref["bbb_step_model_W0"], rtol=3e-7, atol=6e-8)


# ---------------------------------------------------------------------------------------
# Interval bound propagation over sampled weights
# ---------------------------------------------------------------------------------------
def _mlp_engine(dims, oracle, seed=4, sigma_scale=0.5):
    import deepbayes_b200 as db
    from deepbayes_b200.engine import Engine
    L = oracle.mlp(dims)
    layers = [db.Dense(dims[1], activation="relu", input_shape=(dims[0],))]
    for h in dims[2:-1]:
        layers.append(db.Dense(h, activation="relu"))
    layers.append(db.Dense(dims[-1], activation="softmax"))
    eng = Engine(db.Sequential(layers))
    mean, std = oracle.synthetic_posterior(L, seed=seed, sigma_scale=sigma_scale)
    eng.set_gaussian(mean, std)
    return L, eng, mean, std


@pytest.mark.parametrize("dims,B,n", [([30, 24, 16, 7], 5, 6), ([784, 128, 10], 9, 4), ([100, 72, 40, 10], 70, 3)])
def test_ibp_fp32_matches_oracle(oracle, dims, B, n):
    """dbx_ibp_predict (fp32 mode, injected weight noise) against oracle.ibp_forward / chernoff_verification,
    the float64 restatement pinned on the reference's own verifiers.IBP run."""
    L, eng, mean, std = _mlp_engine(dims, oracle)
    x = oracle.synthetic_x((B, dims[0]), seed=1)
    labels = np.random.Generator(np.random.Philox(3)).integers(0, dims[-1], size=B).astype(np.int32)
    noise = oracle.synthetic_eps(L, n, seed=2)
    nf = np.stack([oracle.flatten_list(e) for e in noise])
    eps = 0.01
    out = eng.ibp_predict(x, eps, labels, n, noise=nf, mode="fp32", want=("worst", "lower", "upper", "counts"), return_flags=True)
    want_l = np.zeros((B, dims[-1])); want_u = np.zeros((B, dims[-1]))
    for s in range(n):
        w = oracle.sample_gaussian(mean, std, noise[s], order="f32")
        lo, hi = oracle.ibp_forward(L, w, x, eps)
        want_l += lo; want_u += hi
    scale = max(np.abs(want_l).max(), np.abs(want_u).max())
    np.testing.assert_allclose(out["lower"].cpu().numpy(), want_l, rtol=1e-5, atol=2e-6 * scale)
    np.testing.assert_allclose(out["upper"].cpu().numpy(), want_u, rtol=1e-5, atol=2e-6 * scale)
    total, flags = oracle.chernoff_verification(L, mean, std, x, eps, labels, n, noise, order="f32")
    np.testing.assert_allclose(out["worst"].cpu().numpy(), total, rtol=2e-5, atol=1e-6)
    np.testing.assert_array_equal(out["flags"].cpu().numpy().astype(bool), flags)
    np.testing.assert_array_equal(out["counts"].cpu().numpy(), flags.sum(0))
    assert (out["lower"] <= out["upper"]).all()


def test_ibp_matches_reference_run(oracle, golden_dir):
    """The CUDA path against the committed outputs of the reference's own verifiers.IBP and
    chernoff_bound_verification on the VOGN tutorial posterior (tests/golden/reference_ibp.npz)."""
    import deepbayes_b200 as db
    from deepbayes_b200 import formats
    from deepbayes_b200.analyzers import verifiers
    ref = np.load(os.path.join(golden_dir, "reference_ibp.npz"))
    pm = db.PosteriorModel(os.path.join(golden_dir, "posteriors", "VOGN_MNIST_Posterior"), mode="fp32")
    x, eps = ref["x"], float(ref["eps"][0])
    lo, hi = verifiers.IBP(pm, x, pm.posterior_mean, eps)
    np.testing.assert_allclose(lo, ref["ibp_mean_l"], rtol=1e-5, atol=2e-5)
    np.testing.assert_allclose(hi, ref["ibp_mean_u"], rtol=1e-5, atol=2e-5)
    # chernoff_bound_verification: the reference's noise (global NumPy RNG, seed 424242) injected
    n, cls = int(ref["chernoff_n"][0]), int(ref["chernoff_cls"][0])
    mean = formats.load_tensor_list(os.path.join(pm.path_to_model, "mean.npy"))
    np.random.seed(424242)
    noise = np.stack([np.concatenate([np.random.standard_normal(t.shape).ravel() for t in mean]) for _ in range(n)]).astype(np.float32)
    out = verifiers._engine_of(pm).ibp_predict(x[:1], eps, [cls], n, noise=noise, mode="fp32", want=("worst",))
    np.testing.assert_allclose(out["worst"].cpu().numpy(), ref["chernoff_sum"], rtol=1e-4, atol=1e-5)
    # API shape and sample count of the drop-in function (device Philox noise: same distribution, other draws)
    got = verifiers.chernoff_bound_verification(pm, x[:1], eps, cls, confidence=0.75, delta=0.3)
    assert got.shape == (1, 10) and abs(got.sum() - n) < 1e-3
    # predict=True (regression branch): un-clipped box, last activation on both bounds -- against the oracle restatement
    lo_p, hi_p = verifiers.IBP(pm, x, pm.posterior_mean, eps, predict=True)
    wl, wu = oracle.ibp_forward(oracle.mlp([784, 128, 10]), [np.asarray(t, np.float32) for t in pm.posterior_mean], x, eps, predict=True)
    np.testing.assert_allclose(lo_p, wl, rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(hi_p, wu, rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(lo_p, ref["ibp_mean_predict_l"], rtol=1e-5, atol=2e-6)     # the reference's own run
    np.testing.assert_allclose(hi_p, ref["ibp_mean_predict_u"], rtol=1e-5, atol=2e-6)


def test_ibp_properties_and_bf16(oracle, dbx_opt):
    """eps = 0 collapses the interval onto the plain forward; the bf16 tensor-core back end agrees with the
    bf16 CUDA-core back end and stays within the bf16 tolerance of fp32; stored-sample source; Massart loop."""
    dims, B, n = [784, 256, 256, 10], 200, 12
    L, eng, mean, std = _mlp_engine(dims, oracle, sigma_scale=0.3)
    x = oracle.synthetic_x((B, dims[0]), seed=6)
    labels = np.random.Generator(np.random.Philox(8)).integers(0, 10, size=B).astype(np.int32)
    z = eng.ibp_predict(x, 0.0, labels, n, seed=5, mode="fp32", want=("lower", "upper"))
    np.testing.assert_allclose(z["lower"].cpu().numpy(), z["upper"].cpu().numpy(), rtol=0, atol=1e-4)
    lg, _ = eng.predict_sums(x, n, seed=5, mode="fp32", out_kind="logits")
    np.testing.assert_allclose(z["lower"].cpu().numpy(), lg.cpu().numpy(), rtol=1e-4, atol=1e-3)
    eps = 0.004
    f = eng.ibp_predict(x, eps, labels, n, seed=5, mode="fp32", want=("worst", "lower", "upper", "counts"))
    dbx_opt("NO_TC", "1")
    c = eng.ibp_predict(x, eps, labels, n, seed=5, mode="bf16", want=("worst", "lower", "upper", "counts"))
    assert eng.last_path() == "generic"
    dbx_opt("NO_TC", "0")
    t = eng.ibp_predict(x, eps, labels, n, seed=5, mode="bf16", want=("worst", "lower", "upper", "counts"))
    assert eng.last_path() == "generic_tc"
    for k in ("lower", "upper"):
        np.testing.assert_allclose(t[k].cpu().numpy() / n, c[k].cpu().numpy() / n, rtol=0, atol=3e-3)
        np.testing.assert_allclose(t[k].cpu().numpy() / n, f[k].cpu().numpy() / n, rtol=0, atol=8e-2)
    np.testing.assert_allclose(t["worst"].cpu().numpy() / n, f["worst"].cpu().numpy() / n, rtol=0, atol=BF16_ATOL)
    assert (t["lower"] <= t["upper"]).all()
    assert np.abs(t["counts"].cpu().numpy() - f["counts"].cpu().numpy()).max() <= 2
    # wider ball -> wider interval
    f2 = eng.ibp_predict(x, 2 * eps, labels, n, seed=5, mode="fp32", want=("lower", "upper"))
    assert (f2["lower"] <= f["lower"] + 1e-4).all() and (f["upper"] <= f2["upper"] + 1e-4).all()
    # stored-sample source: rows of a slab drawn with the same ids give the same bounds
    slab = eng.sample(n, seed=5, s0=0)
    eng.set_samples(slab)
    s = eng.ibp_predict(x, eps, labels, n, mode="fp32", stored=True, want=("lower", "upper"))
    np.testing.assert_allclose(s["lower"].cpu().numpy(), f["lower"].cpu().numpy(), rtol=1e-6, atol=1e-5)


@pytest.mark.parametrize("dims,B,n", [([784, 256, 256, 10], 200, 12), ([784, 512, 10], 128, 300), ([104, 72, 200, 136, 10], 130, 5),
                                        ([64, 1024, 10], 1, 3)])
def test_ibp_fused_layer_is_bit_identical_to_two_gemms(oracle, dbx_opt, dims, B, n):
    """ibp_tc_kernel (centre, radius and the interval combine of a hidden layer in one launch, |W| formed in the SM) against
    the two dense_tc launches + ibp_combine_kernel it replaces: same MMA order, same epilogue arithmetic -> same bits.
    Ragged M / N / K tails and the 256-sample chunk boundary (n = 300) included."""
    L, eng, mean, std = _mlp_engine(dims, oracle, sigma_scale=0.3)
    x = oracle.synthetic_x((B, dims[0]), seed=16)
    labels = np.random.Generator(np.random.Philox(18)).integers(0, dims[-1], size=B).astype(np.int32)
    want = ("worst", "lower", "upper", "counts")
    dbx_opt("NO_IBP_TC", "1")
    a = eng.ibp_predict(x, 0.004, labels, n, seed=9, mode="bf16", want=want)
    dbx_opt("NO_IBP_TC", "0")
    k0 = _lib_launches()
    b = eng.ibp_predict(x, 0.004, labels, n, seed=9, mode="bf16", want=want)
    assert _lib_launches() - k0 > 0
    assert eng.last_path() == "generic_tc"
    for k in want:
        np.testing.assert_array_equal(a[k].cpu().numpy(), b[k].cpu().numpy())


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_ibp_through_conv_layers_matches_oracle(oracle, mode):
    """propagate_conv2d (verifiers.py:11-21) + pooling applied per bound (:78-81) + Dense on the device against
    oracle.ibp_forward (whose conv part is pinned on a run of the reference's own function): sampled weights injected."""
    import deepbayes_b200 as db
    from deepbayes_b200.engine import Engine
    m = db.Sequential([db.Conv2D(8, 3, activation="relu", input_shape=(12, 10, 2)), db.Conv2D(6, (2, 3), activation="tanh"),
                       db.MaxPooling2D(2), db.Flatten(), db.Dense(24, activation="relu"), db.Dense(5, activation="softmax")])
    L = [oracle.Conv2D(3, 3, 2, 8, "relu"), oracle.Conv2D(2, 3, 8, 6, "tanh"), oracle.MaxPool2D((2, 2)), oracle.Flatten(),
         oracle.Dense(4 * 3 * 6, 24, "relu"), oracle.Dense(24, 5, "softmax")]
    eng = Engine(m)
    mean, std = oracle.synthetic_posterior(L, seed=5, sigma_scale=0.3)
    eng.set_gaussian(mean, std)
    B, n, eps = 7, 3, 0.01
    x = oracle.synthetic_x((B, 12, 10, 2), seed=2)
    labels = np.random.Generator(np.random.Philox(3)).integers(0, 5, size=B).astype(np.int32)
    noise = oracle.synthetic_eps(L, n, seed=2)
    nf = np.stack([oracle.flatten_list(e) for e in noise])
    out = eng.ibp_predict(x, eps, labels, n, noise=nf, mode=mode, want=("worst", "lower", "upper", "counts"))
    want_l = np.zeros((B, 5)); want_u = np.zeros((B, 5))
    for s in range(n):
        w = oracle.sample_gaussian(mean, std, noise[s], order="f32")
        lo, hi = oracle.ibp_forward(L, w, x, eps)
        want_l += lo; want_u += hi
    scale = max(np.abs(want_l).max(), np.abs(want_u).max())
    tol = dict(rtol=1e-5, atol=3e-6 * scale) if mode == "fp32" else dict(rtol=0, atol=3e-2 * scale)
    np.testing.assert_allclose(out["lower"].cpu().numpy(), want_l, **tol)
    np.testing.assert_allclose(out["upper"].cpu().numpy(), want_u, **tol)
    assert (out["lower"] <= out["upper"]).all()
    if mode == "fp32":
        # explicit weights through the drop-in function; SAME padding is refused loudly (the reference's call ignores it)
        lo1, hi1 = eng.ibp_one(x.reshape(B, -1), eps, oracle.flatten_list(mean), mode="fp32")
        wl, wu = oracle.ibp_forward(L, mean, x, eps)
        np.testing.assert_allclose(lo1.cpu().numpy(), wl, rtol=1e-5, atol=3e-6 * scale)
        np.testing.assert_allclose(hi1.cpu().numpy(), wu, rtol=1e-5, atol=3e-6 * scale)
        m2 = db.Sequential([db.Conv2D(4, 3, padding="same", activation="relu", input_shape=(6, 6, 1)), db.Flatten(), db.Dense(3, activation="softmax")])
        e2 = Engine(m2)
        e2.set_gaussian(np.zeros(e2.P, np.float32), np.ones(e2.P, np.float32))
        with pytest.raises(Exception, match="valid"):
            e2.ibp_predict(np.zeros((1, 6, 6, 1), np.float32), 0.1, [0], 2, mode="fp32", want=("worst",))


def test_bf16_sampler_with_tensors_at_odd_offsets(oracle):
    """Kernel tensors whose flat Keras offset is not a multiple of 4 (bias vectors of 5 / 6 entries in front of them): the
    bf16 sampler's 8-byte store path must not be taken for them (it was: misaligned address).  bf16 vs fp32 predict."""
    import deepbayes_b200 as db
    from deepbayes_b200.engine import Engine
    m = db.Sequential([db.Conv2D(6, 3, activation="relu", input_shape=(9, 9, 2)), db.Flatten(), db.Dense(24, activation="relu"),
                       db.Dense(5, activation="relu"), db.Dense(8, activation="softmax")])
    eng = Engine(m)
    g = np.random.Generator(np.random.Philox(5))
    eng.set_gaussian((g.standard_normal(eng.P) * 0.1).astype(np.float32), np.full(eng.P, 0.02, np.float32))
    x = oracle.synthetic_x((11, 9, 9, 2), seed=3)
    a, _ = eng.predict_sums(x, 6, seed=2, mode="bf16")
    b, _ = eng.predict_sums(x, 6, seed=2, mode="fp32")
    np.testing.assert_allclose(a.cpu().numpy() / 6, b.cpu().numpy() / 6, rtol=0, atol=BF16_ATOL)


def test_massart_with_ibp(tmp_path, oracle):
    import deepbayes_b200 as db
    from deepbayes_b200.analyzers import verifiers
    L = oracle.mlp([784, 128, 10])
    m = db.Sequential([db.Dense(128, activation="relu", input_shape=(784,)), db.Dense(10, activation="softmax")])
    opt = db.optimizers.BayesByBackprop().compile(m, None, batch_size=16, inflate_prior=1.0)
    x = oracle.synthetic_x((1, 784), seed=9)
    cls = int(np.argmax(opt.predict(x)))
    p, it, _ = verifiers.massart_bound_check(opt, x, 1e-4, None, cls=cls, verify=True, classification=True, confidence=0.8, delta=0.3)
    assert 0.0 <= p <= 1.0 and 1 <= it <= verifiers.chernoff_sample_bound(0.8, 0.3) + 1
    # a Python predicate (verifiers.py:628) on every sample's worst case == the device predicate on the same sample ids
    opt._next_sample = 0
    seen = []
    pa = verifiers.massart_bound_check(opt, x, 1e-4, None, cls=cls, classification=True, confidence=0.8, delta=0.3)
    opt._next_sample = 0
    pb = verifiers.massart_bound_check(opt, x, 1e-4, lambda wc: (seen.append(np.asarray(wc)), np.argmax(np.squeeze(wc)) == cls)[1],
                                       cls=cls, classification=True, confidence=0.8, delta=0.3)
    assert pa[:2] == pb[:2]
    assert len(seen) == int(pb[1]) and seen[0].shape == (1, 10) and np.allclose(seen[0].sum(), 1, rtol=1e-5)
    with pytest.raises(ValueError):   # no predicate and nothing to build the device predicate from
        verifiers.massart_bound_check(opt, x, 1e-4, None, classification=True)


def test_massart_regression_branch_and_chernoff_override_vs_oracle(tmp_path, oracle):
    """verifiers.py:609-619 (verify=True, classification=False: IBP(predict=True) on the un-clipped box, per output the bound
    farther from ``cls``) and the chernoff=True override (:585, :643-648) walked against the oracle on the SAME posterior
    samples (Philox ids 35.. -- the first 35 go to the `mean = 0.0 * model.predict(inp)` line, :583)."""
    import deepbayes_b200 as db
    from deepbayes_b200 import formats
    from deepbayes_b200.analyzers import verifiers
    dims = [12, 16, 3]
    L = oracle.mlp(dims, out_act="linear")
    m = db.Sequential([db.Dense(16, activation="relu", input_shape=(12,)), db.Dense(3, activation="linear")])
    mean, std = oracle.synthetic_posterior(L, seed=4, sigma_scale=0.3)
    pdir = str(tmp_path / "post")
    formats.save_gaussian_posterior(pdir, m, mean, std, {"input_upper": 1.0, "input_lower": 0.0})
    seed = 21
    pm = db.PosteriorModel(pdir, mode="fp32", seed=seed)
    x = oracle.synthetic_x((1, 12), seed=3)
    eps = 0.05
    target = oracle.forward(L, mean, x, out="logits").reshape(-1)
    tol = 0.35
    pred = lambda wc: bool(np.abs(np.squeeze(wc) - target).max() < tol)
    seen = []
    frac, iters, mz = verifiers.massart_bound_check(pm, x, eps, lambda wc: (seen.append(np.array(wc)), pred(wc))[1], cls=target,
                                                    confidence=0.8, delta=0.3)
    P = oracle.param_count(L)
    outcomes, worsts = [], []
    for i in range(verifiers.chernoff_sample_bound(0.8, 0.3) + 2):
        w = oracle.sample_gaussian(mean, std, oracle.split_flat(L, oracle.philox_normals(seed, 35 + i, P)))
        lo, hi = oracle.ibp_forward(L, w, x, eps, predict=True)
        wc = np.where(np.abs(hi - target) > np.abs(lo - target), hi, lo)
        worsts.append(wc); outcomes.append(pred(wc))
    want_frac, want_iters = oracle.massart_loop(np.asarray(outcomes, np.uint8), confidence=0.8, delta=0.3)
    for got, want in zip(seen, worsts):
        np.testing.assert_allclose(np.squeeze(got), np.squeeze(want), rtol=2e-5, atol=2e-5)
    margins = [abs(np.abs(np.squeeze(wc) - target).max() - tol) for wc in worsts[:int(want_iters)]]
    if min(margins) > 1e-3:
        assert (frac, iters) == (want_frac, want_iters)
    assert np.all(np.asarray(mz) == 0)
    assert 0.0 < want_frac < 1.0 or len(set(outcomes)) == 1
    # chernoff=True (:585, :643-648) on a classifier: the walk runs to the Chernoff bound whatever the halting rule says, and the
    # mean of model._predict(glob_adv) over the iterations comes back (rows of softmax outputs: they sum to 1)
    Lc = oracle.mlp([12, 16, 3])
    mc = db.Sequential([db.Dense(16, activation="relu", input_shape=(12,)), db.Dense(3, activation="softmax")])
    cdir = str(tmp_path / "post_c")
    formats.save_gaussian_posterior(cdir, mc, mean, std, {"input_upper": 1.0, "input_lower": 0.0})
    pm2 = db.PosteriorModel(cdir, mode="fp32", seed=seed)
    c0 = int(np.argmax(pm2.predict(x)))
    f2, it2, mean2 = verifiers.massart_bound_check(pm2, x, eps, lambda wc: np.argmax(np.squeeze(wc)) == c0, cls=c0, confidence=0.8,
                                                   delta=0.3, chernoff=True, verify=False)
    assert it2 == verifiers.chernoff_sample_bound(0.8, 0.3) + 1
    assert np.asarray(mean2).shape == (1, 3) and np.allclose(np.asarray(mean2).sum(), 1.0, rtol=1e-4)


@pytest.mark.parametrize("dims,B,S", [([784, 1024, 1024, 10], 64, 20), ([200, 512, 384, 12], 150, 9), ([784, 640, 10], 40, 7)])
def test_stored_posterior_f32_slab_on_tensor_cores(oracle, dbx_opt, dims, B, S):
    """Stored posteriors at B > 16 in bf16 mode: dense_tc_f32w_kernel reads the Dense kernels from the fp32 slab by TMA and
    rounds them to bf16 inside the SM.  Same arithmetic as the two-pass path (bf16 copy of the slab chunk, then
    dense_tc_kernel), so the two agree to accumulation order; both stay within the bf16 tolerance of the fp32 mode."""
    L, eng, mean, std = _mlp_engine(dims, oracle, sigma_scale=0.5)
    slab = eng.sample(S, seed=3, s0=0)
    eng.set_samples(slab)
    x = oracle.synthetic_x((B, dims[0]), seed=2)
    g = np.random.Generator(np.random.Philox(5))
    idx = g.integers(0, S, size=2 * S).astype(np.int32)
    wts = g.uniform(0.5, 1.5, size=2 * S).astype(np.float32)
    f1, f2 = eng.predict_stored_sums(x, 2 * S, idx=idx, wts=wts, mode="fp32", second_moment=True)
    t1, t2 = eng.predict_stored_sums(x, 2 * S, idx=idx, wts=wts, mode="bf16", second_moment=True)
    assert eng.last_path() == "generic_tc"
    dbx_opt("NO_F32W", "1")
    c1, c2 = eng.predict_stored_sums(x, 2 * S, idx=idx, wts=wts, mode="bf16", second_moment=True)
    dbx_opt("NO_F32W", "0")
    tot = float(wts.sum())
    np.testing.assert_allclose(t1.cpu().numpy() / tot, c1.cpu().numpy() / tot, rtol=0, atol=2e-4)
    np.testing.assert_allclose(t2.cpu().numpy() / tot, c2.cpu().numpy() / tot, rtol=0, atol=2e-4)
    np.testing.assert_allclose(t1.cpu().numpy() / tot, f1.cpu().numpy() / tot, rtol=0, atol=BF16_ATOL)
    assert (t1.argmax(1) == f1.argmax(1)).float().mean().item() > 0.97
    # contiguous rows j0.. (no index list)
    a1, _ = eng.predict_stored_sums(x, S - 2, j0=2, mode="bf16")
    dbx_opt("NO_F32W", "1")
    b1, _ = eng.predict_stored_sums(x, S - 2, j0=2, mode="bf16")
    np.testing.assert_allclose(a1.cpu().numpy() / (S - 2), b1.cpu().numpy() / (S - 2), rtol=0, atol=2e-4)


def test_direct_conv_on_tensor_cores(oracle, dbx_opt):
    """conv_tc_kernel (per-tap TMA boxes of the NHWC input, no im2col matrix) against the im2col + GEMM path on the same
    sampled weights, and against the fp32 mode: VALID and SAME padding, Cin = 16 / 32 / 64, ragged row groups, Cout
    that is not a multiple of 64."""
    import deepbayes_b200 as db
    from deepbayes_b200.engine import Engine
    m = db.Sequential([db.Conv2D(16, 3, activation="relu", input_shape=(13, 11, 3)),          # Cin = 3: im2col path
                       db.Conv2D(32, 3, padding="same", activation="relu"),                   # Cin = 16, SAME
                       db.Conv2D(64, (2, 3), activation="tanh"),                              # Cin = 32, VALID, 2x3 taps
                       db.Conv2D(24, 3, padding="same", activation="relu"),                   # Cin = 64, Cout = 24
                       db.MaxPooling2D(2), db.Flatten(), db.Dense(10, activation="softmax")])
    eng = Engine(m)
    g = np.random.Generator(np.random.Philox(21))
    mu = (g.standard_normal(eng.P) * 0.08).astype(np.float32)
    sg = np.full(eng.P, 0.01, np.float32)
    eng.set_gaussian(mu, sg)
    B, n = 37, 5
    x = oracle.synthetic_x((B, 13, 11, 3), seed=4)
    f1, _ = eng.predict_sums(x, n, seed=9, mode="fp32")
    d1, d2 = eng.predict_sums(x, n, seed=9, mode="bf16", second_moment=True)
    assert eng.last_path() == "generic_tc"
    launches_direct = _lib_launches()
    dbx_opt("NO_DIRECT_CONV", "1")
    i1, i2 = eng.predict_sums(x, n, seed=9, mode="bf16", second_moment=True)
    np.testing.assert_allclose(d1.cpu().numpy() / n, i1.cpu().numpy() / n, rtol=0, atol=1.5e-3)
    np.testing.assert_allclose(d2.cpu().numpy() / n, i2.cpu().numpy() / n, rtol=0, atol=1.5e-3)
    np.testing.assert_allclose(d1.cpu().numpy() / n, f1.cpu().numpy() / n, rtol=0, atol=BF16_ATOL)
    assert _lib_launches() - launches_direct > 0
    # the persistent GEMM kernel (normally only for >= 296 tiles) on every GEMM of this model
    dbx_opt("PERSIST_MIN_TILES", "1")
    p1, p2 = eng.predict_sums(x, n, seed=9, mode="bf16", second_moment=True)
    dbx_opt("NO_PERSIST", "1")
    q1, q2 = eng.predict_sums(x, n, seed=9, mode="bf16", second_moment=True)
    np.testing.assert_allclose(p1.cpu().numpy() / n, q1.cpu().numpy() / n, rtol=0, atol=2e-4)
    np.testing.assert_allclose(p2.cpu().numpy() / n, q2.cpu().numpy() / n, rtol=0, atol=2e-4)
    np.testing.assert_allclose(p1.cpu().numpy() / n, i1.cpu().numpy() / n, rtol=0, atol=1.5e-3)


@pytest.mark.parametrize("shape,c1,c2,pad", [((13, 11, 3), 16, 32, "valid"), ((20, 18, 3), 32, 64, "same"), ((9, 30, 3), 16, 8, "valid"),
                                            ((32, 32, 3), 32, 64, "valid")])
def test_conv_with_fused_maxpool_is_bit_identical(oracle, dbx_opt, shape, c1, c2, pad):
    """conv_tc_kernel with the following 2x2 / stride-2 MaxPool done in its epilogue (staged tile in shared memory, packed
    bf16x2 max, pooled NHWC rows out) against the same kernel + maxpool_bf16_vec8_kernel: same bf16 values before the max,
    so the same bits after it.  Odd output heights / widths (the last row / column is dropped), SAME padding, Cout = 8 .. 64."""
    import deepbayes_b200 as db
    from deepbayes_b200.engine import Engine
    m = db.Sequential([db.Conv2D(c1, 3, activation="relu", input_shape=shape), db.Conv2D(c2, 3, padding=pad, activation="relu"),
                       db.MaxPooling2D(2), db.Flatten(), db.Dense(16, activation="relu"), db.Dense(10, activation="softmax")])
    eng = Engine(m)
    g = np.random.Generator(np.random.Philox(23))
    eng.set_gaussian((g.standard_normal(eng.P) * 0.08).astype(np.float32), np.full(eng.P, 0.01, np.float32))
    B, n = 19, 30
    x = oracle.synthetic_x((B,) + shape, seed=4)
    dbx_opt("NO_POOL_FUSE", "1")
    a1, a2 = eng.predict_sums(x, n, seed=9, mode="bf16", second_moment=True)
    k0 = _lib_launches()
    dbx_opt("NO_POOL_FUSE", "0")
    b1, b2 = eng.predict_sums(x, n, seed=9, mode="bf16", second_moment=True)
    k1 = _lib_launches()
    dbx_opt("NO_POOL_FUSE", "1")
    eng.predict_sums(x, n, seed=9, mode="bf16", second_moment=True)
    k2 = _lib_launches()
    assert k1 - k0 < k2 - k1          # the pooling launches are gone
    np.testing.assert_array_equal(a1.cpu().numpy(), b1.cpu().numpy())
    np.testing.assert_array_equal(a2.cpu().numpy(), b2.cpu().numpy())
    f1, _ = eng.predict_sums(x, n, seed=9, mode="fp32")
    np.testing.assert_allclose(b1.cpu().numpy() / n, f1.cpu().numpy() / n, rtol=0, atol=BF16_ATOL)


@pytest.mark.parametrize("dims,B,n", [([784, 1024, 1024, 10], 600, 40), ([512, 640, 10], 300, 70), ([256, 1000, 328, 10], 257, 90)])
def test_pair_gemm_is_bit_identical_to_single_cta_gemm(oracle, dbx_opt, dims, B, n):
    """dense_tc_pair_kernel (cta_group::2, 256 x 256 tiles, each CTA staging half of the weight block) against the single-CTA
    tensor-core GEMMs on the same sampled weights: same products in the same k order -> same bits.  Ragged row / column
    tiles (600 = 2 x 256 + 88 rows, 640 = 2.5 x 256 columns, 1000 and 328 columns) included."""
    L, eng, mean, std = _mlp_engine(dims, oracle, sigma_scale=0.3)
    x = oracle.synthetic_x((B, dims[0]), seed=21)
    dbx_opt("NO_FUSED", "1")
    dbx_opt("NO_PAIR_GEMM", "1")
    a1, a2 = eng.predict_sums(x, n, seed=3, mode="bf16", second_moment=True)
    assert eng.last_path() == "generic_tc"
    dbx_opt("NO_PAIR_GEMM", "0")
    b1, b2 = eng.predict_sums(x, n, seed=3, mode="bf16", second_moment=True)
    np.testing.assert_array_equal(a1.cpu().numpy(), b1.cpu().numpy())
    np.testing.assert_array_equal(a2.cpu().numpy(), b2.cpu().numpy())
    f1, _ = eng.predict_sums(x, n, seed=3, mode="fp32")
    np.testing.assert_allclose(b1.cpu().numpy() / n, f1.cpu().numpy() / n, rtol=0, atol=BF16_ATOL)


def _lib_launches():
    from deepbayes_b200 import _lib
    return _lib.launch_count()


# ---------------------------------------------------------------------------------------
# Bayes-by-Backprop training step
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("dims,B", [([12, 9, 7, 5], 6), ([784, 128, 10], 64), ([100, 72, 40, 10], 37)])
def test_bbb_step_matches_oracle(oracle, dims, B):
    """dbx_bbb_step with injected noise against oracle.bbb_step (hand-written gradients, checked by finite differences
    of the loss that is pinned on a run of the reference's losses.py): updated mean / rho, loss, KL component,
    predictions, sampled weights."""
    import deepbayes_b200 as db
    L = oracle.mlp(dims)
    layers = [db.Dense(dims[1], activation="relu", input_shape=(dims[0],))]
    for h in dims[2:-1]:
        layers.append(db.Dense(h, activation="relu"))
    layers.append(db.Dense(dims[-1], activation="softmax"))
    opt = db.optimizers.BayesByBackprop().compile(db.Sequential(layers), None, batch_size=B, inflate_prior=2.0, kl_weight=0.7)
    mean, std = oracle.synthetic_posterior(L, seed=1, sigma_scale=3.0)
    opt.posterior_mean = [m.copy() for m in mean]
    opt.posterior_var = [oracle.inverse_softplus(s) for s in std]
    rho0 = [r.copy() for r in opt.posterior_var]
    x = oracle.synthetic_x((B, dims[0]), seed=0)
    lab = np.random.Generator(np.random.Philox(3)).integers(0, dims[-1], size=B)
    noise = oracle.synthetic_eps(L, 1, seed=2)[0]
    lr = 0.1
    nm, nr = opt.step(x, lab, lr, noise=noise)
    wm, wr, loss, klc, pred, W = oracle.bbb_step(L, mean, rho0, opt.prior_mean, opt.prior_var, x, lab, noise, lr, kl_weight=0.7)
    for i in range(len(nm)):
        # the step sizes are what is being compared: tolerance relative to the size of the update
        dm, dr = np.abs(wm[i] - mean[i]).max() + 1e-12, np.abs(wr[i] - rho0[i]).max() + 1e-12
        np.testing.assert_allclose(nm[i], wm[i], rtol=0, atol=2e-4 * dm + 1e-7)
        np.testing.assert_allclose(nr[i], wr[i], rtol=0, atol=2e-4 * dr + 1e-7)
        np.testing.assert_allclose(opt.model.get_weights()[i], W[i], rtol=1e-6, atol=1e-7)
    got = opt._last_loss.cpu().numpy()
    np.testing.assert_allclose(got[0], loss, rtol=2e-5)
    np.testing.assert_allclose(got[1], klc, rtol=2e-5)
    np.testing.assert_allclose(opt._last_probs.cpu().numpy(), pred, rtol=2e-5, atol=1e-7)
    with pytest.raises(NotImplementedError):
        db.optimizers.BayesByBackprop().compile(db.Sequential(layers), None, robust_train=5).step(x, lab, lr)


@pytest.mark.parametrize("dims,B,rt", [([784, 512, 512, 10], 256, 0), ([784, 128, 10], 130, 0), ([104, 72, 40, 16], 64, 0),
                                       ([784, 256, 128, 10], 300, 2)])
def test_bbb_step_bf16_on_tensor_cores_matches_oracle(oracle, dims, B, rt):
    """dbx_bbb_step in bf16 mode (compile(..., train_mode="bf16")): forward = dense_tc_kernel, backward-data =
    dense_tc_bwd_kernel, weight gradient = dense_tc_wgrad_kernel (A^T dZ with both operands MN-major), everything else fp32 --
    against the oracle with the same operand roundings and against the exact (fp64) step; the sampled weights, the loss terms
    and the clean predictions are the fp32 step's to rounding."""
    import deepbayes_b200 as db
    L = oracle.mlp(dims)
    layers = [db.Dense(dims[1], activation="relu", input_shape=(dims[0],))]
    for h in dims[2:-1]:
        layers.append(db.Dense(h, activation="relu"))
    layers.append(db.Dense(dims[-1], activation="softmax"))
    kw = dict(robust_train=rt, epsilon=0.02, rob_lam=0.5, linear_schedule=False) if rt else {}
    opt = db.optimizers.BayesByBackprop().compile(db.Sequential(layers), None, batch_size=B, inflate_prior=2.0, kl_weight=0.7, train_mode="bf16", **kw)
    mean, std = oracle.synthetic_posterior(L, seed=1, sigma_scale=3.0)
    opt.posterior_mean = [m.copy() for m in mean]
    opt.posterior_var = [oracle.inverse_softplus(s) for s in std]
    rho0 = [r.copy() for r in opt.posterior_var]
    x = oracle.synthetic_x((B, dims[0]), seed=0)
    lab = np.random.Generator(np.random.Philox(3)).integers(0, dims[-1], size=B)
    noise = oracle.synthetic_eps(L, 1, seed=2)[0]
    lr = 0.1
    nm, nr = opt.step(x, lab, lr, noise=noise)
    assert opt.engine().last_path() == "generic_tc"
    okw = dict(robust_train=rt, epsilon=0.02, robust_lambda=0.5) if rt else {}
    wm16, wr16, loss16, klc16, pred16, W16 = oracle.bbb_step(L, mean, rho0, opt.prior_mean, opt.prior_var, x, lab, noise, lr, kl_weight=0.7,
                                                            dtype=np.float32, bf16=True, **okw)
    wm, wr, loss, klc, pred, W = oracle.bbb_step(L, mean, rho0, opt.prior_mean, opt.prior_var, x, lab, noise, lr, kl_weight=0.7, **okw)
    tol16, tol64 = (2e-2, 0.1) if rt == 0 else (0.15, 0.25)      # (the FGSM sign of near-zero gradient entries moves whole inputs by 2 eps)
    for i in range(len(nm)):
        for got, w16, w64, old in ((nm[i], wm16[i], wm[i], mean[i]), (nr[i], wr16[i], wr[i], rho0[i])):
            d16, d64, dg = (w16 - old).ravel().astype(np.float64), (w64 - old).ravel().astype(np.float64), (got - old).ravel().astype(np.float64)
            assert np.linalg.norm(dg - d16) <= tol16 * np.linalg.norm(d16) + 1e-9, (i, np.linalg.norm(dg - d16) / np.linalg.norm(d16))
            assert np.linalg.norm(dg - d64) <= tol64 * np.linalg.norm(d64) + 1e-9, (i, np.linalg.norm(dg - d64) / np.linalg.norm(d64))
        np.testing.assert_allclose(opt.model.get_weights()[i], W[i], rtol=1e-6, atol=1e-7)        # the draw itself is fp32
    got = opt._last_loss.cpu().numpy()
    np.testing.assert_allclose(got[1], klc, rtol=2e-5)                                              # KL terms: fp32
    np.testing.assert_allclose(got[0], loss16, rtol=5e-3 if rt == 0 else 5e-2)
    if rt == 0:
        np.testing.assert_allclose(opt._last_probs.cpu().numpy(), pred16, rtol=0, atol=1e-2)      # (a stored activation across a bf16 tie)
        # (this posterior is 30x wider than the synthetic posterior of BASELINE, whose bf16 bound on probabilities is BF16_ATOL)
        np.testing.assert_allclose(opt._last_probs.cpu().numpy(), pred, rtol=0, atol=2.5 * BF16_ATOL)


def test_bbb_train_epoch_graph_replay_equals_step_by_step(oracle, dbx_opt):
    """dbx_bbb_train_epoch in bf16 mode replays ONE captured step as a CUDA graph for the full mini-batches (step id, first row and
    output slot come from a device-side context the graph advances): parameters, per-step losses, predictions and the last
    sampled weights must be bit-identical to launching every step's kernels one by one (the default; option TRAIN_GRAPH=1
    selects the replay, which measured slower), ragged last mini-batch included."""
    import torch
    import deepbayes_b200 as db
    dims, bs, N = [784, 256, 128, 10], 256, 6 * 256 + 100
    L = oracle.mlp(dims)
    layers = [db.Dense(256, activation="relu", input_shape=(784,)), db.Dense(128, activation="relu"), db.Dense(10, activation="softmax")]
    X = oracle.synthetic_x((N, 784), seed=5)
    y = np.random.Generator(np.random.Philox(6)).integers(0, 10, size=N).astype(np.int32)
    res = []
    for graph in (1, 0):
        dbx_opt("TRAIN_GRAPH", graph)
        opt = db.optimizers.BayesByBackprop().compile(db.Sequential(layers), None, batch_size=bs, inflate_prior=1.0, seed=3, train_mode="bf16")
        mean, std = oracle.synthetic_posterior(L, seed=1, sigma_scale=1.0)
        eng = opt.engine()
        dm = eng._dev_f32(oracle.flatten_list(mean)).clone()
        dr = eng._dev_f32(oracle.flatten_list([oracle.inverse_softplus(s_) for s_ in std])).clone()
        pm_ = eng._dev_f32(oracle.flatten_list(opt.prior_mean)).clone()
        pv_ = eng._dev_f32(oracle.flatten_list(opt.prior_var)).clone()
        Xd, yd = torch.from_numpy(X).to(eng.device), torch.from_numpy(y).to(eng.device)
        out = eng.bbb_train_epoch(Xd, yd, bs, dm, dr, pm_, pv_, 0.05, 0.5, seed=11, step0=40, mode="bf16")
        assert eng.last_path() == "generic_tc"
        res.append([t.cpu().numpy() for t in (dm, dr, out["loss"], out["probs"], out["W"])])
    for a, b in zip(*res):
        np.testing.assert_array_equal(a, b)
    assert np.isfinite(res[0][2]).all() and res[0][2].shape == (7, 2)


def test_bbb_training_in_bf16_mode_learns(oracle):
    """A few epochs of BayesByBackprop.train with train_mode="bf16" (one dbx_bbb_train_epoch call per epoch on the tensor-core
    step) reach the accuracy of the fp32 run on the same data."""
    import deepbayes_b200 as db
    g = np.random.Generator(np.random.Philox(11))
    C, D, N = 4, 24, 2048
    centers = g.standard_normal((C, D)).astype(np.float32) * 2.0
    y = g.integers(0, C, size=N)
    X = (centers[y] + g.standard_normal((N, D)).astype(np.float32) * 0.5).astype(np.float32)
    accs = {}
    for tm in ("fp32", "bf16"):
        m = db.Sequential([db.Dense(32, activation="relu", input_shape=(D,)), db.Dense(C, activation="softmax")])
        opt = db.optimizers.BayesByBackprop().compile(m, None, batch_size=256, learning_rate=0.3, epochs=5, inflate_prior=1.0, kl_weight=0.1, seed=5,
                                                      train_mode=tm)
        opt.train(X[:1536], y[:1536], X[1536:], y[1536:])
        assert opt.loss_log[-1] < opt.loss_log[0]
        accs[tm] = (opt.acc_log[-1], opt.val_log[-1][1])
    assert accs["bf16"][0] > 0.9 and accs["bf16"][1] > 0.85 and abs(accs["bf16"][0] - accs["fp32"][0]) < 0.05


def test_bbb_training_learns(oracle):
    """End to end: a few epochs of BayesByBackprop.train on a separable synthetic problem drive the loss down and the
    accuracy up, and the trained posterior predicts through the fused / generic inference path."""
    import deepbayes_b200 as db
    g = np.random.Generator(np.random.Philox(11))
    C, D, N = 4, 20, 2048
    centers = g.standard_normal((C, D)).astype(np.float32) * 2.0
    y = g.integers(0, C, size=N)
    X = (centers[y] + g.standard_normal((N, D)).astype(np.float32) * 0.5).astype(np.float32)
    m = db.Sequential([db.Dense(32, activation="relu", input_shape=(D,)), db.Dense(C, activation="softmax")])
    opt = db.optimizers.BayesByBackprop().compile(m, None, batch_size=128, learning_rate=0.3, epochs=5, inflate_prior=1.0, kl_weight=0.1, seed=5)
    opt.train(X[:1536], y[:1536], X[1536:], y[1536:])
    assert opt.loss_log[-1] < opt.loss_log[0] and opt.acc_log[-1] > 0.9 and opt.val_log[-1][1] > 0.85
    pm = np.asarray(opt.engine().predict_sums(X[1536:], 16, seed=1, mode="fp32")[0].cpu().numpy()) / 16
    assert (pm.argmax(1) == y[1536:]).mean() > 0.85


@pytest.mark.parametrize("robust_train", [0, 2])
def test_bbb_epoch_on_device_equals_per_step_loop(oracle, robust_train):
    """`dbx_bbb_train_epoch` (one device call per epoch, training set resident in HBM) against the per-mini-batch loop of
    `step` calls (optimizer.py:112-116): the same kernels with the same Philox step ids in the same order, so the trained
    posterior is bit-identical; ragged last mini-batch included."""
    import deepbayes_b200 as db
    g = np.random.Generator(np.random.Philox(12))
    C, D, N = 5, 24, 1000
    centers = g.standard_normal((C, D)).astype(np.float32)
    y = g.integers(0, C, size=N)
    X = np.clip(0.5 + 0.2 * centers[y] + 0.1 * g.standard_normal((N, D)), 0, 1).astype(np.float32)
    res = []
    for dev in (True, False):
        m = db.Sequential([db.Dense(48, activation="relu", input_shape=(D,)), db.Dense(C, activation="softmax")])
        if res:
            m.set_weights(res[0]["w0"])
        w0 = [np.array(w) for w in m.get_weights()]
        opt = db.optimizers.BayesByBackprop().compile(m, None, batch_size=96, learning_rate=0.2, epochs=2, inflate_prior=1.0, kl_weight=0.1,
                                                      seed=7, robust_train=robust_train, epsilon=0.02, rob_lam=0.5, linear_schedule=False)
        opt.train(X, y, device_epochs=dev)
        res.append({"w0": w0, "mean": opt.posterior_mean, "var": opt.posterior_var, "loss": list(opt.loss_log), "acc": list(opt.acc_log),
                    "W": opt.model.get_weights()})
    a, b = res
    for i in range(len(a["mean"])):
        np.testing.assert_array_equal(np.asarray(a["mean"][i]), np.asarray(b["mean"][i]))
        np.testing.assert_array_equal(np.asarray(a["var"][i]), np.asarray(b["var"][i]))
        np.testing.assert_array_equal(np.asarray(a["W"][i]), np.asarray(b["W"][i]))
    np.testing.assert_allclose(a["loss"], b["loss"], rtol=1e-5)
    np.testing.assert_allclose(a["acc"], b["acc"], rtol=0, atol=1e-7)


@pytest.mark.parametrize("dims,B,n_outer,n_inner", [([10, 8, 6, 4], 3, 2, 5), ([784, 128, 10], 40, 1, 35), ([100, 72, 40, 10], 33, 3, 1)])
def test_input_gradient_matches_oracle(oracle, dims, B, n_outer, n_inner):
    """dbx_input_gradient (injected weight noise) against oracle.input_gradient, summed over the outer iterations."""
    L, eng, mean, std = _mlp_engine(dims, oracle, seed=2, sigma_scale=2.0)
    x = oracle.synthetic_x((B, dims[0]), seed=4)
    lab = np.random.Generator(np.random.Philox(9)).integers(0, dims[-1], size=B)
    noise = oracle.synthetic_eps(L, n_outer * n_inner, seed=3)
    nf = np.stack([oracle.flatten_list(e) for e in noise])
    got = eng.input_gradient(x, lab, n_outer, n_inner, noise=nf).cpu().numpy()
    want = np.zeros((B, dims[0]))
    for it in range(n_outer):
        ws = [oracle.sample_gaussian(mean, std, noise[it * n_inner + s], order="f32") for s in range(n_inner)]
        want += oracle.input_gradient(L, ws, x, lab)
    np.testing.assert_allclose(got, want, rtol=2e-4, atol=2e-6 * np.abs(want).max())
    # one explicit weight vector == one stored sample
    w = oracle.flatten_list(oracle.sample_gaussian(mean, std, noise[0], order="f32")).astype(np.float32)
    g1 = eng.input_gradient_one(x, lab, w).cpu().numpy()
    np.testing.assert_allclose(g1, oracle.input_gradient(L, [oracle.split_flat(L, w)], x, lab), rtol=2e-4, atol=2e-6 * np.abs(g1).max() + 1e-12)


@pytest.mark.parametrize("dims,B,n_outer,n_inner", [([784, 128, 10], 40, 1, 35), ([784, 512, 512, 10], 130, 2, 4), ([104, 72, 40, 16], 33, 3, 1),
                                                    ([784, 512, 512, 10], 300, 1, 3)])
@pytest.mark.parametrize("persist", [False, True])
def test_input_gradient_bf16_on_tensor_cores_matches_oracle(oracle, dbx_opt, dims, B, n_outer, n_inner, persist):
    """dbx_input_gradient in bf16 mode (forward = per-sample tcgen05 GEMMs, backward = dense_tc_bwd_kernel: dZ * W_s^T with the
    K-major view of the Keras kernel, activation derivative in its epilogue) against the oracle with the same roundings
    (operands to bf16, fp32 accumulation) and against the fp64 oracle of attacks.py:49-75.  persist: every GEMM launch through
    the persistent kernels (dense_tc_persist_kernel / dense_tc_bwd_persist_kernel), otherwise one CTA per tile at these sizes."""
    dbx_opt("PERSIST_MIN_TILES", "1" if persist else "1000000")
    L, eng, mean, std = _mlp_engine(dims, oracle, seed=2, sigma_scale=2.0)
    x = oracle.synthetic_x((B, dims[0]), seed=4)
    lab = np.random.Generator(np.random.Philox(9)).integers(0, dims[-1], size=B)
    noise = oracle.synthetic_eps(L, n_outer * n_inner, seed=3)
    nf = np.stack([oracle.flatten_list(e) for e in noise])
    got = eng.input_gradient(x, lab, n_outer, n_inner, noise=nf, mode="bf16").cpu().numpy()
    assert eng.last_path() == "generic_tc"
    want16 = np.zeros((B, dims[0]))
    want = np.zeros((B, dims[0]))
    for it in range(n_outer):
        ws = [oracle.sample_gaussian(mean, std, noise[it * n_inner + s], order="f32") for s in range(n_inner)]
        want16 += oracle.input_gradient(L, ws, x, lab, bf16=True)
        want += oracle.input_gradient(L, ws, x, lab)
    scale = np.abs(want).max()
    # same roundings: what is left is fp32 summation order, which moves a stored activation across a bf16 tie or a relu
    # across zero in a few rows (their gradient then differs by that unit's share); most rows agree to fp32 rounding
    row_err = np.abs(got - want16).max(1)
    assert np.median(row_err) <= 1e-3 * scale and np.quantile(row_err, 0.75) <= 4e-3 * scale and row_err.max() <= 0.1 * scale, \
        (np.median(row_err) / scale, np.quantile(row_err, 0.75) / scale, row_err.max() / scale)
    # against the exact gradient: bf16 operand rounding (2^-9 relative per operand, through two or three layers each way)
    # (observed: 2-6 % of the gradient's norm at this posterior width)
    assert np.linalg.norm(got - want) <= 0.1 * np.linalg.norm(want)
    np.testing.assert_allclose(got, want, rtol=0, atol=0.25 * scale)
    big = np.abs(want) > 0.5 * scale
    assert (np.sign(got[big]) == np.sign(want[big])).all()
    # the fp32 kernels on the same inputs
    g32 = eng.input_gradient(x, lab, n_outer, n_inner, noise=nf, mode="fp32").cpu().numpy()
    assert eng.last_path() == "generic"
    np.testing.assert_allclose(g32, want, rtol=2e-4, atol=1e-5 * scale)


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_input_gradient_over_stored_samples_drawn_with_replacement(oracle, mode):
    """gradient_expectation over a sample-based posterior: each forward of predict() uses a stored sample drawn with
    replacement (posteriormodel.py:77, :95-97) -- the rows are given as an index list."""
    dims, B, S = [784, 128, 64, 10], 20, 7
    L, eng, mean, std = _mlp_engine(dims, oracle, seed=2, sigma_scale=2.0)
    noise = oracle.synthetic_eps(L, S, seed=3)
    samples = [oracle.sample_gaussian(mean, std, noise[s], order="f32") for s in range(S)]
    eng.set_samples(np.stack([oracle.flatten_list(w) for w in samples]).astype(np.float32))
    x = oracle.synthetic_x((B, dims[0]), seed=4)
    lab = np.random.Generator(np.random.Philox(9)).integers(0, dims[-1], size=B)
    idx = np.array([3, 3, 0, 6, 1, 3, 5, 2, 6, 6])
    got = eng.input_gradient(x, lab, 2, 5, stored=True, idx=idx, mode=mode).cpu().numpy()
    want = sum(oracle.input_gradient(L, [samples[j] for j in idx[it * 5:it * 5 + 5]], x, lab, bf16=(mode == "bf16")) for it in range(2))
    scale = np.abs(want).max()
    if mode == "fp32":
        np.testing.assert_allclose(got, want, rtol=2e-4, atol=2e-6 * scale)
    else:
        assert eng.last_path() == "generic_tc"
        row_err = np.abs(got - want).max(1)
        assert np.median(row_err) <= 1e-3 * scale and row_err.max() <= 0.1 * scale
    with pytest.raises(ValueError):
        eng.input_gradient(x, lab, 2, 5, stored=True, idx=[0, 1, 2, 3, 4, 5, 6, 7, 0, 1], mode=mode)      # row 7 does not exist


@pytest.mark.parametrize("persist", [False, True])
def test_count_predicate_fgsm_bf16_on_tensor_cores(oracle, dbx_opt, persist):
    """dbx_count_predicate_fgsm / dbx_fgsm_samples in bf16 mode (forward, backward, second forward at every sample's own
    adversarial input on tcgen05) against the oracle's per-sample FGSM walk; entries whose gradient sign or top-2 margin is
    within bf16 rounding of a tie are left out."""
    dims, B, n = [784, 256, 128, 10], 70, 9
    dbx_opt("PERSIST_MIN_TILES", "1" if persist else "1000000")
    L, eng, mean, std = _mlp_engine(dims, oracle, seed=2, sigma_scale=2.0)
    x = oracle.synthetic_x((B, dims[0]), seed=4)
    g = np.random.Generator(np.random.Philox(9))
    direc = g.integers(0, dims[-1], size=B)
    lab = g.integers(0, dims[-1], size=B)
    noise = oracle.synthetic_eps(L, n, seed=3)
    nf = np.stack([oracle.flatten_list(e) for e in noise])
    eps = 0.03
    counts, flags = eng.count_predicate_fgsm(x, direc, lab, eps, n, noise=nf, lower=0.0, upper=1.0, return_flags=True, mode="bf16")
    assert eng.last_path() == "generic_tc"
    probs = eng.fgsm_samples(x, direc, eps, n, noise=nf, lower=0.0, upper=1.0, mode="bf16").cpu().numpy()
    want = np.zeros((n, B), bool)
    ok = np.ones((n, B), bool)
    wprobs = np.zeros((n, B, dims[-1]))
    for s in range(n):
        w = oracle.sample_gaussian(mean, std, noise[s], order="f32")
        gr = oracle.input_gradient(L, [w], x, direc, bf16=True)
        adv = oracle.fgsm(oracle.bf16_round(x), gr, eps, 0.0, 1.0)
        z = oracle.forward(L, w, adv, out="logits", bf16=True)
        wprobs[s] = oracle.forward(L, w, adv, bf16=True)
        want[s] = z.argmax(1) == lab
        top2 = np.sort(z, axis=1)[:, -2:]
        # a sign flip of a small gradient entry moves that input by 2 eps: rows where many entries are near zero are not comparable
        near_zero = (np.abs(gr) < 2e-2 * np.abs(gr).max(1, keepdims=True)).mean(1)
        ok[s] = ((top2[:, 1] - top2[:, 0]) > 0.05 * np.abs(z).max()) & (near_zero < 0.5)
    got = flags.cpu().numpy().astype(bool)
    assert ok.mean() > 0.5 and (got == want)[ok].mean() > 0.97
    np.testing.assert_array_equal(counts.cpu().numpy(), got.sum(0))
    np.testing.assert_array_equal(probs.argmax(2) == lab[None, :], got)
    assert np.median(np.abs(probs - wprobs).max(2)) < 2e-2


def test_fgsm_and_pgd_lower_the_true_class_probability(oracle, golden_dir):
    """FGSM / PGD through the drop-in API on the VOGN tutorial posterior: the adversarial inputs stay in the eps-ball and in
    [input_lower, input_upper], and the posterior-mean probability of the attacked class goes down."""
    import deepbayes_b200 as db
    from deepbayes_b200 import analyzers
    pm = db.PosteriorModel(os.path.join(golden_dir, "posteriors", "VOGN_MNIST_Posterior"), mode="fp32")
    x = oracle.synthetic_x((16, 784), seed=12)
    base = np.asarray(pm.predict(x, n=35)).reshape(16, 10)
    cls = base.argmax(1)
    eps = 0.05
    adv = analyzers.FGSM(pm, x, None, eps, samples=3)
    assert adv.shape == x.shape and (np.abs(adv - x) <= eps + 1e-6).all()
    pa = np.asarray(pm.predict(adv, n=35)).reshape(16, 10)
    assert pa[np.arange(16), cls].mean() < base[np.arange(16), cls].mean() - 1e-3
    np.random.seed(0)
    adv2 = analyzers.PGD(pm, x, None, eps, num_steps=3, num_models=2)
    assert (np.abs(adv2 - x) <= eps + 1e-6).all()
    p2 = np.asarray(pm.predict(adv2, n=35)).reshape(16, 10)
    assert p2[np.arange(16), cls].mean() < base[np.arange(16), cls].mean() - 1e-3
    with pytest.raises(NotImplementedError):
        analyzers.PGD(pm, x, None, eps, num_steps=1, order=0)      # the reference's _PGD builds a gradient for order=1 only (:152-155)
    # CW (attacks.py:179-226): Adam-style sign steps on the same expected gradient
    adv3 = analyzers.CW(pm, x, None, eps, num_steps=4, num_models=2)
    assert adv3.shape == x.shape and (np.abs(adv3 - x) <= eps + 1e-6).all()
    p3 = np.asarray(pm.predict(adv3, n=35)).reshape(16, 10)
    assert p3[np.arange(16), cls].mean() < base[np.arange(16), cls].mean() - 1e-3
    with pytest.raises(NotImplementedError):
        analyzers.CW(pm, x, None, eps, order=0)


def test_zeroth_order_gradient_and_fgsm_order0(oracle):
    """attacks.py:17-45: the central-difference gradient through model.predict alone.  On an optimizer object (predict = one forward
    with the current weights) it must agree with the analytic input gradient of the same weights; a CNN goes through as well
    (no backward kernels needed), and FGSM(order=0) lowers the attacked class's probability."""
    import deepbayes_b200 as db
    from deepbayes_b200 import analyzers
    dims = [24, 32, 5]
    L = oracle.mlp(dims, hidden_act="tanh")      # smooth: central differences are second-order accurate everywhere
    m = db.Sequential([db.Dense(32, activation="tanh", input_shape=(24,)), db.Dense(5, activation="softmax")])
    mean, _ = oracle.synthetic_posterior(L, seed=3, sigma_scale=1.0)
    m.set_weights([w * 3.0 for w in mean])
    opt = db.optimizers.BayesByBackprop().compile(m, None, batch_size=8, inflate_prior=1.0)
    x = oracle.synthetic_x((3, 24), seed=4)
    lab = np.array([1, 4, 0])
    g0 = analyzers.zeroth_order_gradient(opt, x, lab, None, num_models=1, step=0.01)
    assert g0.shape == x.shape
    W = [np.asarray(w) for w in opt.model.get_weights()]
    for i in range(3):      # the reference differences the PER-IMAGE loss: batch of one
        want = oracle.input_gradient(L, [W], x[i:i + 1], lab[i:i + 1])[0]
        np.testing.assert_allclose(g0[i], want, rtol=0, atol=0.02 * np.abs(want).max() + 1e-4)
    base = np.asarray(opt.predict(x)).reshape(3, 5)
    cls = base.argmax(1)
    adv = analyzers.FGSM(opt, x, None, 0.1, order=0, samples=1)
    assert adv.shape == x.shape and (np.abs(adv - x) <= 0.1 + 1e-6).all()
    pa = np.asarray(opt.predict(adv)).reshape(3, 5)
    assert pa[np.arange(3), cls].mean() < base[np.arange(3), cls].mean()
    # a convolutional model: order 0 needs the forward path only
    cnn = db.Sequential([db.Conv2D(4, 3, activation="relu", input_shape=(6, 6, 1)), db.Flatten(), db.Dense(3, activation="softmax")])
    copt = db.optimizers.BayesByBackprop().compile(cnn, None, batch_size=4, inflate_prior=1.0)
    xc = oracle.synthetic_x((2, 6, 6, 1), seed=5)
    gc = analyzers.zeroth_order_gradient(copt, xc, np.array([0, 2]), None, num_models=1, step=0.05)
    assert gc.shape == xc.shape and np.isfinite(gc).all() and np.abs(gc).max() > 0


def test_count_predicate_fgsm_matches_oracle(oracle):
    """dbx_count_predicate_fgsm (injected weight noise): per posterior sample, the FGSM step of attacks.py:94-101 against that
    sample's weights and the predicate at the adversarial input, against the oracle's input_gradient / fgsm / forward."""
    dims, B, n = [30, 24, 16, 7], 9, 11
    L, eng, mean, std = _mlp_engine(dims, oracle, seed=2, sigma_scale=2.0)
    x = oracle.synthetic_x((B, dims[0]), seed=4)
    g = np.random.Generator(np.random.Philox(9))
    direc = g.integers(0, dims[-1], size=B)
    lab = g.integers(0, dims[-1], size=B)
    noise = oracle.synthetic_eps(L, n, seed=3)
    nf = np.stack([oracle.flatten_list(e) for e in noise])
    eps = 0.03
    counts, flags = eng.count_predicate_fgsm(x, direc, lab, eps, n, noise=nf, lower=0.0, upper=1.0, return_flags=True)
    want = np.zeros((n, B), bool)
    margin_ok = np.ones((n, B), bool)
    for s in range(n):
        w = oracle.sample_gaussian(mean, std, noise[s], order="f32")
        gr = oracle.input_gradient(L, [w], x, direc)
        adv = oracle.fgsm(x, gr, eps, 0.0, 1.0)
        z = oracle.forward(L, w, adv, dtype=np.float64, out="logits")
        want[s] = z.argmax(1) == lab
        top2 = np.sort(z, axis=1)[:, -2:]
        margin_ok[s] = ((top2[:, 1] - top2[:, 0]) > 1e-4) & (np.abs(gr).min(1) > 1e-9)   # ties / zero gradients may flip in fp32
    got = flags.cpu().numpy().astype(bool)
    assert (got == want)[margin_ok].all() and margin_ok.mean() > 0.9
    np.testing.assert_array_equal(counts.cpu().numpy(), got.sum(0))


def test_massart_with_device_fgsm(oracle):
    import deepbayes_b200 as db
    from deepbayes_b200.analyzers import verifiers
    m = db.Sequential([db.Dense(64, activation="relu", input_shape=(784,)), db.Dense(10, activation="softmax")])
    opt = db.optimizers.BayesByBackprop().compile(m, None, batch_size=16, inflate_prior=1.0)
    x = oracle.synthetic_x((1, 784), seed=9)
    cls = int(np.argmax(opt.predict(x)))
    p, it, _ = verifiers.massart_bound_check(opt, x, 0.01, None, cls=cls, verify=False, confidence=0.8, delta=0.3)
    assert 0.0 <= p <= 1.0 and 1 <= it <= verifiers.chernoff_sample_bound(0.8, 0.3) + 1


def test_bbb_step_fgsm_adversarial_training_matches_oracle(oracle):
    """robust_train = 2 (bayesbybackprop.py:91-101): FGSM against the step's own weights, loss on the mix of clean and
    adversarial predictions, both backward passes -- device vs oracle with injected noise."""
    import deepbayes_b200 as db
    dims, B = [40, 24, 16, 6], 21
    L = oracle.mlp(dims)
    layers = [db.Dense(24, activation="relu", input_shape=(40,)), db.Dense(16, activation="relu"), db.Dense(6, activation="softmax")]
    opt = db.optimizers.BayesByBackprop().compile(db.Sequential(layers), None, batch_size=B, inflate_prior=2.0, kl_weight=0.5,
                                                  robust_train=2, epsilon=0.05, rob_lam=0.3)
    mean, std = oracle.synthetic_posterior(L, seed=1, sigma_scale=3.0)
    opt.posterior_mean = [m.copy() for m in mean]
    opt.posterior_var = [oracle.inverse_softplus(s) for s in std]
    rho0 = [r.copy() for r in opt.posterior_var]
    x = oracle.synthetic_x((B, 40), seed=0)
    lab = np.random.Generator(np.random.Philox(3)).integers(0, 6, size=B)
    noise = oracle.synthetic_eps(L, 1, seed=2)[0]
    lr = 0.1
    nm, nr = opt.step(x, lab, lr, noise=noise)
    wm, wr, loss, klc, pred, W = oracle.bbb_step(L, mean, rho0, opt.prior_mean, opt.prior_var, x, lab, noise, lr, kl_weight=0.5,
                                                 robust_train=2, epsilon=0.05, robust_lambda=0.3)
    for i in range(len(nm)):
        dm, dr = np.abs(wm[i] - mean[i]).max() + 1e-12, np.abs(wr[i] - rho0[i]).max() + 1e-12
        np.testing.assert_allclose(nm[i], wm[i], rtol=0, atol=5e-4 * dm + 1e-7)
        np.testing.assert_allclose(nr[i], wr[i], rtol=0, atol=5e-4 * dr + 1e-7)
    got = opt._last_loss.cpu().numpy()
    np.testing.assert_allclose(got[0], loss, rtol=5e-5)
    np.testing.assert_allclose(opt._last_probs.cpu().numpy(), pred, rtol=2e-5, atol=1e-7)


@pytest.mark.parametrize("hidden_act", ["relu", "tanh"])
def test_bbb_step_ibp_training_matches_oracle(oracle, hidden_act):
    """robust_train = 1 (bayesbybackprop.py:73-90): interval forward through the step's own weights, softmax of the
    worst-case logits mixed with the clean prediction, backward through BOTH the plain and the interval forward --
    device vs oracle (float64, gradients checked by finite differences in test_oracle.py) with injected noise."""
    import deepbayes_b200 as db
    dims, B = [40, 24, 16, 6], 21
    L = oracle.mlp(dims, hidden_act=hidden_act)
    layers = [db.Dense(24, activation=hidden_act, input_shape=(40,)), db.Dense(16, activation=hidden_act), db.Dense(6, activation="softmax")]
    opt = db.optimizers.BayesByBackprop().compile(db.Sequential(layers), None, batch_size=B, inflate_prior=2.0, kl_weight=0.5,
                                                  robust_train=1, epsilon=0.02, rob_lam=0.3)
    mean, std = oracle.synthetic_posterior(L, seed=1, sigma_scale=3.0)
    opt.posterior_mean = [m.copy() for m in mean]
    opt.posterior_var = [oracle.inverse_softplus(s) for s in std]
    rho0 = [r.copy() for r in opt.posterior_var]
    x = oracle.synthetic_x((B, 40), seed=0)
    lab = np.random.Generator(np.random.Philox(3)).integers(0, 6, size=B)
    noise = oracle.synthetic_eps(L, 1, seed=2)[0]
    lr = 0.1
    nm, nr = opt.step(x, lab, lr, noise=noise)
    wm, wr, loss, klc, pred, W = oracle.bbb_step(L, mean, rho0, opt.prior_mean, opt.prior_var, x, lab, noise, lr, kl_weight=0.5,
                                                 robust_train=1, epsilon=0.02, robust_lambda=0.3)
    for i in range(len(nm)):
        dm, dr = np.abs(wm[i] - mean[i]).max() + 1e-12, np.abs(wr[i] - rho0[i]).max() + 1e-12
        np.testing.assert_allclose(nm[i], wm[i], rtol=0, atol=5e-4 * dm + 1e-7)
        np.testing.assert_allclose(nr[i], wr[i], rtol=0, atol=5e-4 * dr + 1e-7)
    got = opt._last_loss.cpu().numpy()
    np.testing.assert_allclose(got[0], loss, rtol=5e-5)
    np.testing.assert_allclose(opt._last_probs.cpu().numpy(), pred, rtol=2e-5, atol=1e-7)
    # lambda = 1 switches the interval term off: same update as the standard step
    opt1 = db.optimizers.BayesByBackprop().compile(db.Sequential(layers), None, batch_size=B, inflate_prior=2.0, kl_weight=0.5,
                                                   robust_train=1, epsilon=0.02, rob_lam=1.0, linear_schedule=False)
    opt0 = db.optimizers.BayesByBackprop().compile(db.Sequential(layers), None, batch_size=B, inflate_prior=2.0, kl_weight=0.5)
    for o in (opt0, opt1):
        o.posterior_mean = [m.copy() for m in mean]
        o.posterior_var = [r.copy() for r in rho0]
    if float(getattr(opt1, "robust_lambda", 1.0)) == 1.0:
        a, _ = opt0.step(x, lab, lr, noise=noise)
        b, _ = opt1.step(x, lab, lr, noise=noise)
        for i in range(len(a)):
            np.testing.assert_allclose(a[i], b[i], rtol=0, atol=1e-7)


def test_bbb_step_ibp_mc_and_fgsm_mc_modes(oracle):
    """robust_train = 3 (IBP averaged over injected epsilon draws, `dbx_bbb_step_ibp_mc`) against the oracle, and
    robust_train = 4 (the same FGSM pass loss_mc times, :123-136) against the oracle's mode 2 at lambda = 0."""
    import deepbayes_b200 as db
    dims, B = [40, 24, 16, 6], 21
    L = oracle.mlp(dims)
    layers = [db.Dense(24, activation="relu", input_shape=(40,)), db.Dense(16, activation="relu"), db.Dense(6, activation="softmax")]
    mean, std = oracle.synthetic_posterior(L, seed=1, sigma_scale=3.0)
    x = oracle.synthetic_x((B, 40), seed=0)
    lab = np.random.Generator(np.random.Philox(3)).integers(0, 6, size=B)
    noise = oracle.synthetic_eps(L, 1, seed=2)[0]
    lr, draws = 0.1, [0.03, 0.004, 0.011]
    for rt in (3, 4):
        opt = db.optimizers.BayesByBackprop().compile(db.Sequential(layers), None, batch_size=B, inflate_prior=2.0, kl_weight=0.5,
                                                      robust_train=rt, epsilon=0.05, rob_lam=0.3, loss_mc=3)
        opt.posterior_mean = [m.copy() for m in mean]
        opt.posterior_var = [oracle.inverse_softplus(s) for s in std]
        rho0 = [r.copy() for r in opt.posterior_var]
        nm, nr = opt.step(x, lab, lr, noise=noise, eps_draws=draws if rt == 3 else None)
        if rt == 3:
            want = oracle.bbb_step(L, mean, rho0, opt.prior_mean, opt.prior_var, x, lab, noise, lr, kl_weight=0.5, robust_train=3, eps_draws=draws)
        else:
            want = oracle.bbb_step(L, mean, rho0, opt.prior_mean, opt.prior_var, x, lab, noise, lr, kl_weight=0.5, robust_train=2,
                                   epsilon=0.05, robust_lambda=0.0)
        wm, wr, loss, klc, pred, W = want
        for i in range(len(nm)):
            dm, dr = np.abs(wm[i] - mean[i]).max() + 1e-12, np.abs(wr[i] - rho0[i]).max() + 1e-12
            np.testing.assert_allclose(nm[i], wm[i], rtol=0, atol=5e-4 * dm + 1e-7)
            np.testing.assert_allclose(nr[i], wr[i], rtol=0, atol=5e-4 * dr + 1e-7)
        np.testing.assert_allclose(opt._last_loss.cpu().numpy()[0], loss, rtol=5e-5)
        np.testing.assert_allclose(opt._last_probs.cpu().numpy(), pred, rtol=2e-5, atol=1e-7)
    # un-injected draws: Exponential with mean epsilon from the optimizer's own generator, one device call per step
    opt = db.optimizers.BayesByBackprop().compile(db.Sequential(layers), None, batch_size=B, robust_train=3, epsilon=0.02, loss_mc=2)
    opt.step(x, lab, lr)
    assert np.isfinite(opt._last_loss.cpu().numpy()).all()


def test_bbb_ibp_training_tightens_the_bounds(oracle):
    """A few hundred robust_train = 1 steps on a separable toy problem: the clean accuracy stays high and the certified
    (IBP) worst-case probability of the true class ends above what standard training reaches."""
    import deepbayes_b200 as db
    from deepbayes_b200.analyzers import verifiers
    g = np.random.Generator(np.random.Philox(11))
    n, d = 512, 20
    lab = g.integers(0, 2, size=n)
    x = np.clip(0.5 + 0.25 * (2 * lab[:, None] - 1) * np.ones((1, d)) + 0.05 * g.standard_normal((n, d)), 0, 1).astype(np.float32)
    res = {}
    for rt in (0, 1):
        layers = [db.Dense(32, activation="relu", input_shape=(d,)), db.Dense(2, activation="softmax")]
        opt = db.optimizers.BayesByBackprop().compile(db.Sequential(layers), None, batch_size=128, learning_rate=0.3, epochs=1, inflate_prior=1.0,
                                                      kl_weight=0.01, robust_train=rt, epsilon=0.1, rob_lam=0.5, linear_schedule=False, seed=5)
        for it in range(300):
            sel = g.integers(0, n, size=128)
            opt.step(x[sel], lab[sel], 0.3, sync=(it == 299))
        w = opt.posterior_mean
        acc = (np.argmax(opt.predict(x), 1) == lab).mean()
        lo, hi = verifiers.IBP(opt, x, w, 0.1)
        worst = np.where(np.eye(2)[lab] > 0, lo, hi)
        cert = (np.argmax(worst, 1) == lab).mean()
        res[rt] = (acc, cert)
    assert res[1][0] > 0.95 and res[1][1] >= res[0][1] - 0.02 and res[1][1] > 0.9, res


def test_ibp_state_weight_margin_matches_reference_run(oracle, golden_dir):
    """dbx_ibp_state (explicit input box, weight margin in units of sigma) against the committed output of the
    reference's own verifiers.IBP_state on the VOGN tutorial posterior, and against the oracle on a synthetic model."""
    import deepbayes_b200 as db
    from deepbayes_b200.analyzers import verifiers
    ref = np.load(os.path.join(golden_dir, "reference_ibp.npz"))
    pm = db.PosteriorModel(os.path.join(golden_dir, "posteriors", "VOGN_MNIST_Posterior"), mode="fp32")
    lo, hi = verifiers.IBP_state(pm, ref["state_s0"], ref["state_s1"], pm.posterior_mean, weight_margin=0.5)
    scale = np.abs(ref["state_u"]).max()
    np.testing.assert_allclose(lo, ref["state_l"], rtol=1e-5, atol=3e-6 * scale)
    np.testing.assert_allclose(hi, ref["state_u"], rtol=1e-5, atol=3e-6 * scale)
    assert verifiers.IBP_prob is verifiers.IBP_state
    dims, B = [50, 36, 20, 6], 23
    L, eng, mean, std = _mlp_engine(dims, oracle, seed=7, sigma_scale=1.5)
    x = oracle.synthetic_x((B, 50), seed=3)
    s0, s1 = (x - 0.02).astype(np.float32), (x + 0.03).astype(np.float32)
    w = oracle.flatten_list(mean).astype(np.float32)
    for marg in (0.0, 0.7):
        glo, ghi = eng.ibp_state(s0, s1, w, marg)
        wlo, whi = oracle.ibp_state(L, mean, s0, s1, std, marg)
        sc = max(np.abs(wlo).max(), np.abs(whi).max())
        np.testing.assert_allclose(glo.cpu().numpy(), wlo, rtol=1e-5, atol=3e-6 * sc)
        np.testing.assert_allclose(ghi.cpu().numpy(), whi, rtol=1e-5, atol=3e-6 * sc)
