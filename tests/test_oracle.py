"""CPU tests: the oracle against closed forms, known-answer vectors and the committed golden
outputs ('what the new repo must supply itself')."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import cases  # noqa: E402


def test_philox_known_answers(oracle):
    # Random123 kat_vectors for philox4x32-10
    kat = [([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
           ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
           ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
            [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1])]
    for c, k, want in kat:
        got = oracle.philox4x32_10(np.array([c], np.uint32), np.array([k], np.uint32))[0]
        assert [int(v) for v in got] == want


def test_philox_normals_moments(oracle):
    z = oracle.philox_normals(123, 5, 400000).astype(np.float64)
    assert abs(z.mean()) < 5e-3 and abs(z.std() - 1) < 5e-3
    assert abs((z ** 3).mean()) < 2e-2 and abs((z ** 4).mean() - 3) < 5e-2
    # different sample ids / seeds give different streams
    assert not np.allclose(z[:100], oracle.philox_normals(123, 6, 100))
    assert not np.allclose(z[:100], oracle.philox_normals(124, 5, 100))


def test_softplus_thresholds(oracle):
    x = np.array([-80, -20, -13.95, -13.9, -1, 0, 1, 13.9, 13.95, 20, 100], np.float32)
    y = oracle.softplus(x)
    ref = np.log1p(np.exp(x.astype(np.float64)))
    np.testing.assert_allclose(y, ref, rtol=2e-6, atol=0)
    assert y[-1] == np.float32(100) and y[0] == np.exp(np.float32(-80))
    s = np.array([1e-3, 0.05, 0.5, 2.0], np.float32)
    np.testing.assert_allclose(oracle.softplus(oracle.inverse_softplus(s)), s, rtol=2e-4)


def test_sigma_zero_is_deterministic(oracle):
    L = oracle.mlp([20, 16, 5])
    mean, std = oracle.synthetic_posterior(L, seed=3)
    std0 = [np.zeros_like(s) for s in std]
    x = oracle.synthetic_x((7, 20), seed=4)
    eps = oracle.synthetic_eps(L, 5, seed=5)
    got = oracle.predict(L, mean, std0, x, 5, eps)
    np.testing.assert_allclose(got, oracle.forward(L, mean, x), rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(got.sum(-1), 1, rtol=1e-5)


def test_predict_equals_explicit_loop(oracle):
    L = oracle.mlp([12, 9, 4])
    mean, std = oracle.synthetic_posterior(L, seed=3, sigma_scale=0.5)
    x = oracle.synthetic_x((6, 12), seed=4)
    eps = oracle.synthetic_eps(L, 4, seed=5)
    acc = np.zeros((6, 4), np.float32)
    for e in eps:
        w = [m.astype(np.float64) + s.astype(np.float64) * ee.astype(np.float64) for m, s, ee in zip(mean, std, e)]
        acc += oracle.forward(L, [t.astype(np.float32) for t in w], x)
    np.testing.assert_array_equal(oracle.predict(L, mean, std, x, 4, eps), acc / np.float32(4.0))


def test_fp32_vs_fp64_forward(oracle):
    c = cases.case_inputs("mlp_cfg2")
    w = oracle.sample_gaussian(c["mean"], c["std"], c["eps"][0])
    y32 = oracle.forward(c["layers"], w, c["x"], np.float32)
    y64 = oracle.forward(c["layers"], w, c["x"], np.float64)
    np.testing.assert_allclose(y32, y64, rtol=2e-5, atol=1e-7)


def test_bf16_round(oracle):
    import torch
    a = np.random.Generator(np.random.Philox(0)).standard_normal(10000).astype(np.float32)
    a[:4] = [0.0, -0.0, 1.0000001e-38, 3.3895314e38]
    want = torch.from_numpy(a).to(torch.bfloat16).to(torch.float32).numpy()
    np.testing.assert_array_equal(oracle.bf16_round(a), want)


def test_conv_against_torch(oracle):
    import torch
    g = np.random.Generator(np.random.Philox(1))
    x = g.standard_normal((2, 9, 8, 3)).astype(np.float32)
    for strides, padding in (((1, 1), "valid"), ((2, 2), "same"), ((1, 2), "same")):
        l = oracle.Conv2D(3, 3, 3, 5, "linear", strides, padding)
        k = g.standard_normal((3, 3, 3, 5)).astype(np.float32)
        got = oracle._conv2d(x, k, l)
        xt = torch.from_numpy(x).permute(0, 3, 1, 2)
        kt = torch.from_numpy(k).permute(3, 2, 0, 1)
        if padding == "same":
            ho, wo = -(-9 // strides[0]), -(-8 // strides[1])
            ph = max((ho - 1) * strides[0] + 3 - 9, 0); pw = max((wo - 1) * strides[1] + 3 - 8, 0)
            xt = torch.nn.functional.pad(xt, (pw // 2, pw - pw // 2, ph // 2, ph - ph // 2))
        want = torch.nn.functional.conv2d(xt, kt, stride=strides).permute(0, 2, 3, 1).numpy()
        np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-5)


def test_flops_and_params(oracle):
    assert oracle.param_count(oracle.mlp([784, 512, 10])) == 407050
    assert oracle.param_count(oracle.mlp([784, 512, 512, 10])) == 669706
    assert oracle.flops_per_forward(oracle.mlp([784, 512, 512, 10])) == 1337344
    assert oracle.param_count(oracle.cnn_cfg4()) == 1626442
    assert oracle.param_count(oracle.mlp([784, 1024, 1024, 10])) == 1863690


def test_chernoff_and_massart(oracle):
    assert oracle.chernoff_n(0.95, 0.3) == 380
    assert oracle.chernoff_n(0.995, 0.01) == 105967
    assert abs(oracle.okamoto_bound(0.05, 0.3) - 379.42) < 0.01
    assert oracle.absolute_massart_halting(1, 2, [0.2, 0.8], 0.05, 0.3, 0.05) == -1
    flags = np.ones(1000, np.uint8)
    frac, iters = oracle.massart_loop(flags)
    assert frac == 1.0 and iters < 380  # all-success sequence halts before the Chernoff bound
    lb, ub = oracle.clopper_pearson(0.0, 1.0)
    assert lb == 0.0 and 0.9 < ub <= 1.0


@pytest.mark.parametrize("name", cases.NAMES)
def test_oracle_matches_committed_golden(oracle, golden_dir, name):
    blob = np.load(os.path.join(golden_dir, "oracle_cases.npz"))
    exp = cases.expected(name)
    for k, v in exp.items():
        want = blob["%s/%s" % (name, k)]
        if v.dtype.kind in "iu":
            np.testing.assert_array_equal(v, want)
        else:  # BLAS builds may reorder sums; everything else is bit-stable
            np.testing.assert_allclose(v, want, rtol=2e-5, atol=2e-6, err_msg=k)


def test_fixture_posteriors_load_and_predict(oracle, golden_dir):
    from deepbayes_b200 import formats
    for d in ("VOGN_MNIST_Posterior", "Robust_MNIST_Posterior"):
        p = os.path.join(golden_dir, "posteriors", d)
        mean = formats.load_tensor_list(os.path.join(p, "mean.npy"))
        var = formats.load_tensor_list(os.path.join(p, "var.npy"))
        assert [m.shape for m in mean] == [(784, 128), (128,), (128, 10), (10,)]
        assert all(np.isfinite(v).all() and (v > 0).all() for v in var)
        assert abs(float(var[0].mean()) - 4.666e-4) < 2e-6  # = 1/(N*s), N=60000 (blrvi.py:175)
        info = formats.load_info(p)
        assert info["input_upper"] == np.inf and info["N"] == 60000
        L = oracle.mlp([784, 128, 10])
        x = oracle.synthetic_x((8, 784), seed=0)
        eps = oracle.synthetic_eps(L, 35, seed=2)
        mc = oracle.predict(L, mean, var, x, 35, eps)
        det = oracle.forward(L, mean, x)
        np.testing.assert_allclose(mc.sum(-1), 1.0, rtol=1e-5)
        assert np.abs(mc - det).max() < 2e-3  # sigma ~ 4.7e-4: MC mean hugs the deterministic forward


# ---------------------------------------------------------------------------------------
# The oracle against outputs of the REFERENCE'S OWN source files (executed unmodified on a NumPy
# stand-in for TensorFlow by tests/golden/make_reference_fixtures.py; seeds in that script).
# ---------------------------------------------------------------------------------------
SEED_PM, SEED_BBB = 20240601, 777


def _vogn(golden_dir):
    from deepbayes_b200 import formats
    p = os.path.join(golden_dir, "posteriors", "VOGN_MNIST_Posterior")
    return (formats.load_tensor_list(os.path.join(p, "mean.npy")), formats.load_tensor_list(os.path.join(p, "var.npy")))


def test_legacy_normal_is_loc_plus_scale_times_standard_normal():
    """np.random.normal(loc=array, scale=array) consumes the global stream exactly like
    loc + scale * standard_normal(shape): this is what lets eps be injected."""
    loc = np.linspace(-1, 1, 35).reshape(5, 7).astype(np.float32)
    scale = np.linspace(0.1, 2, 35).reshape(5, 7).astype(np.float32)
    np.random.seed(5)
    a = np.random.normal(loc=loc, scale=scale)
    np.random.seed(5)
    b = loc + scale * np.random.standard_normal(loc.shape)
    np.testing.assert_array_equal(a, b)


def test_oracle_matches_reference_posteriormodel_run(oracle, golden_dir):
    ref = np.load(os.path.join(golden_dir, "reference_run.npz"))
    mean, var = _vogn(golden_dir)
    L = oracle.mlp([784, 128, 10])
    x = oracle.synthetic_x((6, 1, 784), seed=0).astype(np.float64).reshape(6, 784)
    # predict(x, n=7): same global-RNG consumption order as posteriormodel.py:94-101
    np.random.seed(SEED_PM)
    got = oracle.predict(L, mean, var, x, 7, eps=None)
    np.testing.assert_allclose(got.reshape(6, 1, 10), ref["pm_predict_n7"], rtol=5e-6, atol=1e-8)
    # ... which is also what the eps-injection form computes
    np.random.seed(SEED_PM)
    eps = [[np.random.standard_normal(t.shape) for t in mean] for _ in range(7)]
    np.testing.assert_allclose(oracle.predict(L, mean, var, x, 7, eps).reshape(6, 1, 10), ref["pm_predict_n7"], rtol=5e-6, atol=1e-8)
    # predict leaves the LAST sample in the model (posteriormodel.py:96)
    last = oracle.sample_gaussian(mean, var, eps[6])
    np.testing.assert_array_equal(last[2], ref["pm_weights_after_predict_W1"])
    np.random.seed(SEED_PM)
    lg = oracle.predict(L, mean, var, x, 5, eps=None, out="logits")
    np.testing.assert_allclose(lg.reshape(6, 1, 10), ref["pm_predict_logits_n5"], rtol=5e-6, atol=1e-6)
    # sample(inflate=2): var used as a std, float64 result
    np.random.seed(SEED_PM)
    s = oracle.sample_reference_rng(mean, var, inflate=2.0)
    np.testing.assert_array_equal(s[1], ref["pm_sample_inflate2_b0"])
    assert bool(ref["pm_sample_dtype_is_f64"][0]) and s[0].dtype == np.float64
    np.testing.assert_allclose(oracle.forward(L, mean, x).reshape(6, 1, 10), ref["pm_det_predict"], rtol=5e-6, atol=1e-8)  # 3-D vs 2-D matmul path in BLAS
    nb = oracle.forward(L, mean[:-1] + [np.zeros_like(mean[-1])], x, out="logits")
    np.testing.assert_allclose(nb.reshape(6, 1, 10), ref["pm__predict_logits_nobias"], rtol=5e-6, atol=1e-6)


def test_oracle_matches_reference_bbb_run(oracle, golden_dir):
    ref = np.load(os.path.join(golden_dir, "reference_run.npz"))
    L = oracle.mlp([30, 24, 10])
    assert int(ref["bbb_epochs_after_compile"][0]) == 11
    pm, pv = oracle.implicit_prior(L, 2.0)
    rho = [oracle.inverse_softplus(v) for v in pv]
    for i in range(4):
        np.testing.assert_array_equal(pm[i], ref["bbb_prior_mean_%d" % i])
        np.testing.assert_array_equal(pv[i], ref["bbb_prior_var_%d" % i])
        np.testing.assert_allclose(rho[i], ref["bbb_rho_%d" % i], rtol=1e-6)
    noise = [ref["bbb_step_noise_%d" % i] for i in range(4)]
    xb = oracle.synthetic_x((16, 30), seed=3)
    pred, w = oracle.bbb_forward(L, pm, rho, noise, xb)
    np.testing.assert_allclose(pred, ref["bbb_step_predictions"], rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(w[0], ref["bbb_step_model_W0"], rtol=1e-6, atol=1e-8)
    np.random.seed(SEED_BBB)
    s0 = np.random.normal(loc=pm[0], scale=oracle.softplus(rho[0]))
    np.testing.assert_allclose(s0, ref["bbb_sample_W0"], rtol=1e-6, atol=1e-9)


SEED_IBP = 424242


def test_oracle_matches_reference_ibp_run(oracle, golden_dir):
    """oracle.ibp_forward / chernoff_verification against the outputs of the reference's own
    verifiers.IBP / chernoff_bound_verification (tests/golden/make_reference_ibp_fixtures.py)."""
    ref = np.load(os.path.join(golden_dir, "reference_ibp.npz"))
    mean, var = _vogn(golden_dir)
    L = oracle.mlp([784, 128, 10])
    x, eps = ref["x"], float(ref["eps"][0])
    lo, hi = oracle.ibp_forward(L, [np.asarray(t, np.float32) for t in mean], x, eps, predict=False)
    np.testing.assert_allclose(lo, ref["ibp_mean_l"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(hi, ref["ibp_mean_u"], rtol=1e-12, atol=1e-12)
    assert (lo <= hi).all()
    det = oracle.forward(L, mean, x, dtype=np.float64, out="logits")
    assert (lo <= det + 1e-9).all() and (det <= hi + 1e-9).all()        # the interval contains the nominal logits
    lo, hi = oracle.ibp_forward(L, [np.asarray(t, np.float32) for t in mean], x, eps, predict=True)
    np.testing.assert_allclose(lo, ref["ibp_mean_predict_l"], rtol=1e-6, atol=1e-9)   # softmax on the bounds, fp32 in the stand-in
    np.testing.assert_allclose(hi, ref["ibp_mean_predict_u"], rtol=1e-6, atol=1e-9)
    # one sampled weight set: the reference draws with the global RNG, Keras stores fp32
    np.random.seed(SEED_IBP)
    w = [np.asarray(t, np.float32) for t in oracle.sample_reference_rng(mean, var)]
    np.testing.assert_array_equal(w[2], ref["ibp_sample_W1"])
    lo, hi = oracle.ibp_forward(L, w, x, eps)
    np.testing.assert_allclose(lo, ref["ibp_sample_l"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(hi, ref["ibp_sample_u"], rtol=1e-12, atol=1e-12)
    # chernoff_bound_verification: 16 samples, running SUM of softmax(worst case)
    n, cls = int(ref["chernoff_n"][0]), int(ref["chernoff_cls"][0])
    assert n == 16
    np.random.seed(SEED_IBP)
    noise = [[np.random.standard_normal(t.shape) for t in mean] for _ in range(n)]
    total, flags = oracle.chernoff_verification(L, mean, var, x[:1], eps, [cls], n, noise)
    np.testing.assert_allclose(total, ref["chernoff_sum"], rtol=2e-6, atol=1e-7)
    assert flags.shape == (n, 1)


def test_ibp_interval_properties(oracle):
    """Size-independent properties: eps = 0 collapses the interval onto the plain forward; the
    interval widens monotonically with eps; worst-case logits pick lower/upper by class."""
    L = oracle.mlp([30, 24, 16, 7])
    mean, std = oracle.synthetic_posterior(L, seed=4, sigma_scale=0.5)
    x = oracle.synthetic_x((5, 30), seed=1)
    lo0, hi0 = oracle.ibp_forward(L, mean, x, 0.0)
    np.testing.assert_allclose(lo0, hi0, rtol=0, atol=1e-12)
    np.testing.assert_allclose(lo0, oracle.forward(L, mean, x, dtype=np.float64, out="logits"), rtol=1e-6, atol=1e-6)
    lo1, hi1 = oracle.ibp_forward(L, mean, x, 0.01)
    lo2, hi2 = oracle.ibp_forward(L, mean, x, 0.05)
    assert (lo2 <= lo1 + 1e-12).all() and (hi1 <= hi2 + 1e-12).all()
    wc = oracle.worst_case_logits(lo1, hi1, [0, 1, 2, 3, 4])
    for r in range(5):
        assert wc[r, r] == lo1[r, r] and (np.delete(wc[r], r) == np.delete(hi1[r], r)).all()


def test_oracle_matches_reference_kl_loss_run(oracle, golden_dir):
    """oracle.log_q / kl_loss against the reference's own losses.py (log_q, KL_Loss) executed on the stand-in
    (tests/golden/make_reference_bbb_fixtures.py): same inputs, regenerated from seeds."""
    ref = np.load(os.path.join(golden_dir, "reference_bbb.npz"))
    L = oracle.mlp([20, 16, 12, 5])
    mean, std = oracle.synthetic_posterior(L, seed=3, sigma_scale=2.0)
    rho = [oracle.inverse_softplus(s) for s in std]
    pm, pv = oracle.implicit_prior(L, 2.0)
    noise = oracle.synthetic_eps(L, 1, seed=4)[0]
    W = oracle.sample_gaussian(mean, std, noise, order="f32")
    x = oracle.synthetic_x((9, 20), seed=5)
    labels = ref["labels"]
    np.testing.assert_allclose(oracle.log_q(W, mean, rho), ref["logq_post"], rtol=2e-6)
    np.testing.assert_allclose(oracle.log_q(W, pm, pv), ref["logq_prior"], rtol=2e-6)
    for kw in (1.0, 0.25):
        loss, klc, pred = oracle.kl_loss(L, labels, x, W, pm, pv, mean, rho, kw)
        np.testing.assert_allclose(loss, ref["loss_kw%g" % kw], rtol=3e-6)
        np.testing.assert_allclose(klc, ref["klc_kw%g" % kw], rtol=3e-6)
    np.testing.assert_allclose(pred, ref["pred"], rtol=2e-5, atol=1e-7)


def test_bbb_step_gradients_by_finite_differences(oracle):
    """The hand-written gradients of oracle.bbb_step against central differences of the (reference-pinned) loss:
    weight_gradient = dLoss/dW, mean_gradient = dLoss/dmean and var_gradient = dLoss/drho at fixed W -- the three
    tape.gradient calls of bayesbybackprop.py:138-140 -- combined as in :146-163."""
    L = oracle.mlp([12, 9, 7, 5])
    mean, std = oracle.synthetic_posterior(L, seed=1, sigma_scale=3.0)
    rho = [oracle.inverse_softplus(s, np.float64) for s in std]
    pm, pv = oracle.implicit_prior(L, 2.0)
    x = oracle.synthetic_x((6, 12), seed=0)
    lab = np.array([0, 1, 2, 3, 4, 0])
    noise = oracle.synthetic_eps(L, 1, seed=2)[0]
    lr, kw = 0.1, 0.7
    nm, nr, loss, klc, pred, W = oracle.bbb_step(L, mean, rho, pm, pv, x, lab, noise, lr, kl_weight=kw)
    mean64 = [np.asarray(m, np.float64) for m in mean]

    def f(W_, m_, r_):
        return oracle.kl_loss(L, lab, x, W_, pm, pv, m_, r_, kw)[0]
    assert abs(f(W, mean64, rho) - loss) < 1e-12
    g = np.random.Generator(np.random.Philox(0))
    h = 1e-6
    for i in range(len(W)):
        for _ in range(3):
            idx = tuple(g.integers(0, d) for d in W[i].shape)

            def bump(lst, sign):
                out = [np.array(t, np.float64) for t in lst]
                out[i][idx] += sign * h
                return out
            wg = (f(bump(W, 1), mean64, rho) - f(bump(W, -1), mean64, rho)) / (2 * h)
            mg = (f(W, bump(mean64, 1), rho) - f(W, bump(mean64, -1), rho)) / (2 * h)
            vg = (f(W, mean64, bump(rho, 1)) - f(W, mean64, bump(rho, -1))) / (2 * h)
            sig = 1.0 / (1.0 + np.exp(-rho[i][idx]))
            g_mean = (mean64[i][idx] - nm[i][idx]) / lr
            g_rho = (rho[i][idx] - nr[i][idx]) / lr
            np.testing.assert_allclose(g_mean, wg + mg, rtol=2e-5, atol=2e-8)
            np.testing.assert_allclose(g_rho, noise[i][idx] * sig * wg + vg, rtol=2e-5, atol=2e-8)


def test_input_gradient_by_finite_differences(oracle):
    """oracle.input_gradient (hand-written backprop of the mean-prediction cross-entropy, attacks.py:60-67) against central
    differences of oracle.mean_prediction_loss."""
    L = oracle.mlp([10, 8, 6, 4])
    mean, std = oracle.synthetic_posterior(L, seed=2, sigma_scale=2.0)
    noise = oracle.synthetic_eps(L, 5, seed=3)
    ws = [oracle.sample_gaussian(mean, std, e) for e in noise]
    x = oracle.synthetic_x((3, 10), seed=4).astype(np.float64)
    lab = np.array([1, 3, 0])
    g = oracle.input_gradient(L, ws, x, lab)
    h = 1e-6
    for (b, k) in [(0, 0), (1, 7), (2, 9), (0, 5)]:
        xp, xm = x.copy(), x.copy()
        xp[b, k] += h; xm[b, k] -= h
        fd = (oracle.mean_prediction_loss(L, ws, xp, lab)[0] - oracle.mean_prediction_loss(L, ws, xm, lab)[0]) / (2 * h)
        np.testing.assert_allclose(g[b, k], fd, rtol=1e-5, atol=1e-9)
    adv = oracle.fgsm(x, g, 0.05, 0.0, 1.0)
    assert (np.abs(adv - x) <= 0.05 + 1e-12).all() and adv.min() >= 0 and adv.max() <= 1


def test_bbb_step_fgsm_mode_gradient_by_finite_differences(oracle):
    """robust_train = 2: the adversarial input is a constant for the tape (attacks.FGSM returns a NumPy array), so
    dLoss/dW is the derivative of CE(y, lambda f_W(x) + (1 - lambda) f_W(x_adv)) at fixed x_adv."""
    L = oracle.mlp([12, 9, 7, 5])
    mean, std = oracle.synthetic_posterior(L, seed=1, sigma_scale=3.0)
    rho = [oracle.inverse_softplus(s, np.float64) for s in std]
    pm, pv = oracle.implicit_prior(L, 2.0)
    x = oracle.synthetic_x((6, 12), seed=0).astype(np.float64)
    lab = np.array([0, 1, 2, 3, 4, 0])
    noise = oracle.synthetic_eps(L, 1, seed=2)[0]
    lr, kw, lam, eps = 0.1, 0.7, 0.3, 0.05
    nm, nr, loss, klc, pred, W = oracle.bbb_step(L, mean, rho, pm, pv, x, lab, noise, lr, kl_weight=kw, robust_train=2, epsilon=eps, robust_lambda=lam)
    x_adv = oracle.fgsm(x, oracle.input_gradient(L, [W], x, pred.argmax(1)), eps)
    mean64 = [np.asarray(m, np.float64) for m in mean]

    def f(W_, m_, r_):
        out = lam * oracle.forward(L, W_, x, dtype=np.float64) + (1 - lam) * oracle.forward(L, W_, x_adv, dtype=np.float64)
        return oracle.sparse_categorical_crossentropy(lab, out) + kw * (oracle.log_q(W_, m_, r_) - oracle.log_q(W_, pm, pv))
    assert abs(f(W, mean64, rho) - loss) < 1e-12
    g = np.random.Generator(np.random.Philox(1))
    h = 1e-6
    for i in range(len(W)):
        idx = tuple(g.integers(0, d) for d in W[i].shape)

        def bump(lst, sign):
            out = [np.array(t, np.float64) for t in lst]
            out[i][idx] += sign * h
            return out
        wg = (f(bump(W, 1), mean64, rho) - f(bump(W, -1), mean64, rho)) / (2 * h)
        mg = (f(W, bump(mean64, 1), rho) - f(W, bump(mean64, -1), rho)) / (2 * h)
        np.testing.assert_allclose((mean64[i][idx] - nm[i][idx]) / lr, wg + mg, rtol=2e-5, atol=2e-8)


def test_bbb_step_ibp_mode_gradient_by_finite_differences(oracle):
    """robust_train = 1 (bayesbybackprop.py:73-90): dLoss/dW of CE(y, lambda f_W(x) + (1 - lambda) softmax(worst-case IBP
    logits of W)) -- the interval forward is differentiated too (abs(W) -> sign(W), both relu masks)."""
    L = oracle.mlp([12, 9, 7, 5])
    mean, std = oracle.synthetic_posterior(L, seed=1, sigma_scale=3.0)
    rho = [oracle.inverse_softplus(s, np.float64) for s in std]
    pm, pv = oracle.implicit_prior(L, 2.0)
    x = oracle.synthetic_x((6, 12), seed=0).astype(np.float64)
    lab = np.array([0, 1, 2, 3, 4, 0])
    noise = oracle.synthetic_eps(L, 1, seed=2)[0]
    lr, kw, lam, eps = 0.1, 0.7, 0.3, 0.02
    nm, nr, loss, klc, pred, W = oracle.bbb_step(L, mean, rho, pm, pv, x, lab, noise, lr, kl_weight=kw, robust_train=1, epsilon=eps, robust_lambda=lam)
    mean64 = [np.asarray(m, np.float64) for m in mean]

    def f(W_, m_, r_):
        lo, hi = oracle.ibp_forward(L, W_, x, eps)          # the restatement pinned on the reference's verifiers.IBP run
        worst = oracle.worst_case_logits(lo, hi, lab, depth=5)
        e = np.exp(worst - worst.max(1, keepdims=True))
        out = lam * oracle.forward(L, W_, x, dtype=np.float64) + (1 - lam) * e / e.sum(1, keepdims=True)
        return oracle.sparse_categorical_crossentropy(lab, out) + kw * (oracle.log_q(W_, m_, r_) - oracle.log_q(W_, pm, pv))
    assert abs(f(W, mean64, rho) - loss) < 1e-12
    g = np.random.Generator(np.random.Philox(1))
    h = 1e-6
    for i in range(len(W)):
        for _ in range(3):
            idx = tuple(g.integers(0, d) for d in W[i].shape)

            def bump(lst, sign):
                out = [np.array(t, np.float64) for t in lst]
                out[i][idx] += sign * h
                return out
            wg = (f(bump(W, 1), mean64, rho) - f(bump(W, -1), mean64, rho)) / (2 * h)
            mg = (f(W, bump(mean64, 1), rho) - f(W, bump(mean64, -1), rho)) / (2 * h)
            np.testing.assert_allclose((mean64[i][idx] - nm[i][idx]) / lr, wg + mg, rtol=2e-5, atol=2e-8)


def test_bbb_step_ibp_mc_mode_gradient_by_finite_differences(oracle):
    """robust_train = 3 (bayesbybackprop.py:102-121): CE on the mean over epsilon draws of the worst-case softmax."""
    L = oracle.mlp([12, 9, 7, 5])
    mean, std = oracle.synthetic_posterior(L, seed=1, sigma_scale=3.0)
    rho = [oracle.inverse_softplus(s, np.float64) for s in std]
    pm, pv = oracle.implicit_prior(L, 2.0)
    x = oracle.synthetic_x((6, 12), seed=0).astype(np.float64)
    lab = np.array([0, 1, 2, 3, 4, 0])
    noise = oracle.synthetic_eps(L, 1, seed=2)[0]
    lr, kw, draws = 0.1, 0.7, [0.03, 0.004, 0.011]
    nm, nr, loss, klc, pred, W = oracle.bbb_step(L, mean, rho, pm, pv, x, lab, noise, lr, kl_weight=kw, robust_train=3, eps_draws=draws)
    mean64 = [np.asarray(m, np.float64) for m in mean]

    def f(W_, m_, r_):
        out = 0
        for ek in draws:
            lo, hi = oracle.ibp_forward(L, W_, x, ek)
            worst = oracle.worst_case_logits(lo, hi, lab, depth=5)
            e = np.exp(worst - worst.max(1, keepdims=True))
            out = out + e / e.sum(1, keepdims=True) / len(draws)
        return oracle.sparse_categorical_crossentropy(lab, out) + kw * (oracle.log_q(W_, m_, r_) - oracle.log_q(W_, pm, pv))
    assert abs(f(W, mean64, rho) - loss) < 1e-12
    g = np.random.Generator(np.random.Philox(1))
    h = 1e-6
    for i in range(len(W)):
        for _ in range(2):
            idx = tuple(g.integers(0, d) for d in W[i].shape)

            def bump(lst, sign):
                out = [np.array(t, np.float64) for t in lst]
                out[i][idx] += sign * h
                return out
            wg = (f(bump(W, 1), mean64, rho) - f(bump(W, -1), mean64, rho)) / (2 * h)
            mg = (f(W, bump(mean64, 1), rho) - f(W, bump(mean64, -1), rho)) / (2 * h)
            np.testing.assert_allclose((mean64[i][idx] - nm[i][idx]) / lr, wg + mg, rtol=2e-5, atol=2e-8)


def test_oracle_propagate_conv2d_matches_reference_run(oracle, golden_dir):
    """oracle.propagate_conv2d against the reference's own verifiers.propagate_conv2d (:11-21) executed on NumPy stand-ins
    (tests/golden/make_reference_conv_ibp_fixtures.py), and its algebra: the reference ADDS the nominal convolution to both
    interval sums, so centre = 2 conv(mid, W) + b and radius = conv(rad, |W|)."""
    ref = np.load(os.path.join(golden_dir, "reference_conv_ibp.npz"))
    for tag in ("a", "b"):
        W, b, xl, xu = ref[tag + "_W"], ref[tag + "_b"], ref[tag + "_xl"], ref[tag + "_xu"]
        lo, hi = oracle.propagate_conv2d(W, b, xl, xu)
        np.testing.assert_allclose(lo, ref[tag + "_l"], rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(hi, ref[tag + "_u"], rtol=1e-12, atol=1e-12)
        mid, rad = (xl + xu) / 2, (xu - xl) / 2
        np.testing.assert_allclose((hi + lo) / 2, 2 * oracle._conv_valid(mid, W) + b, rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose((hi - lo) / 2, oracle._conv_valid(rad, np.abs(W)), rtol=1e-10, atol=1e-12)
    # IBP through a small conv net: eps = 0 collapses the interval (onto the quirk's doubled-conv network, not the forward)
    L = [oracle.Conv2D(3, 3, 2, 4, "relu"), oracle.MaxPool2D((2, 2)), oracle.Flatten(), oracle.Dense(36, 5, "softmax")]
    mean, _ = oracle.synthetic_posterior(L, seed=3)
    x = oracle.synthetic_x((3, 8, 8, 2), seed=1)
    lo, hi = oracle.ibp_forward(L, mean, x, 0.0)
    np.testing.assert_allclose(lo, hi, rtol=0, atol=1e-12)
    lo2, hi2 = oracle.ibp_forward(L, mean, x, 0.01)
    assert (lo2 <= lo + 1e-12).all() and (hi <= hi2 + 1e-12).all()


def test_oracle_matches_reference_ibp_state_run(oracle, golden_dir):
    """oracle.ibp_state against the reference's own verifiers.IBP_state (weight margin 0.5 sigma, explicit input box)."""
    ref = np.load(os.path.join(golden_dir, "reference_ibp.npz"))
    mean, var = _vogn(golden_dir)
    L = oracle.mlp([784, 128, 10])
    w = [np.asarray(t, np.float32) for t in mean]
    lo, hi = oracle.ibp_state(L, w, ref["state_s0"], ref["state_s1"], var, 0.5)
    np.testing.assert_allclose(lo, ref["state_l"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(hi, ref["state_u"], rtol=1e-12, atol=1e-12)
    lo0, hi0 = oracle.ibp_state(L, w, ref["state_s0"], ref["state_s1"], var, 0.0)
    assert (lo <= lo0 + 1e-12).all() and (hi0 <= hi + 1e-12).all()


def test_bf16_emulating_gradients_stay_close_to_the_exact_ones(oracle):
    """The oracle's bf16 emulation of the widened rows (what the tensor-core gradient loops and training step are compared with on
    the GPU) rounds GEMM operands only: it must stay within bf16 distance of the exact restatement, and be the exact one when no
    rounding can happen (operands already representable: a linear model with bf16-valued weights and inputs)."""
    dims, B = [32, 24, 6], 9
    L = oracle.mlp(dims)
    mean, std = oracle.synthetic_posterior(L, seed=2, sigma_scale=1.0)
    x = oracle.synthetic_x((B, dims[0]), seed=4)
    lab = np.random.Generator(np.random.Philox(9)).integers(0, dims[-1], size=B)
    noise = oracle.synthetic_eps(L, 3, seed=3)
    ws = [oracle.sample_gaussian(mean, std, e, order="f32") for e in noise]
    g64 = oracle.input_gradient(L, ws, x, lab)
    g16 = oracle.input_gradient(L, ws, x, lab, bf16=True)
    assert g16.dtype == np.float32 and 0 < np.linalg.norm(g16 - g64) <= 0.05 * np.linalg.norm(g64)
    # training step: updates of mean within bf16 distance of the exact step, KL component untouched by the emulation
    rho = [oracle.inverse_softplus(s) for s in std]
    pm, pv = oracle.implicit_prior(L, 1.0)
    a = oracle.bbb_step(L, mean, rho, pm, pv, x, lab, noise[0], 0.1, kl_weight=0.5)
    b = oracle.bbb_step(L, mean, rho, pm, pv, x, lab, noise[0], 0.1, kl_weight=0.5, dtype=np.float32, bf16=True)
    for i in range(len(mean)):
        d64, d16 = (a[0][i] - mean[i]).ravel(), (b[0][i] - mean[i]).ravel().astype(np.float64)
        assert np.linalg.norm(d16 - d64) <= 0.05 * np.linalg.norm(d64) + 1e-9
    np.testing.assert_allclose(b[3], a[3], rtol=1e-4)
    with pytest.raises(NotImplementedError):
        oracle.bbb_step(L, mean, rho, pm, pv, x, lab, noise[0], 0.1, robust_train=1, epsilon=0.01, bf16=True)
    # nothing to round: one linear layer, weights and inputs on the bf16 grid -> identical gradients
    L1 = oracle.mlp([16, 4])
    g = np.random.Generator(np.random.Philox(5))
    W1 = [oracle.bf16_round(g.standard_normal((16, 4)).astype(np.float32)), g.standard_normal(4).astype(np.float32)]
    x1 = oracle.bf16_round(g.random((5, 16)).astype(np.float32))
    y1 = g.integers(0, 4, size=5)
    e64, e16 = oracle.input_gradient(L1, [W1], x1, y1), oracle.input_gradient(L1, [W1], x1, y1, bf16=True)
    # (the upstream gradient dz is still rounded to bf16: 2^-9 relative)
    np.testing.assert_allclose(e16, e64, rtol=0, atol=2 ** -8 * np.abs(e64).max())


def test_zeroth_order_attack_matches_a_run_of_the_reference_source(oracle, golden_dir):
    """tests/golden/reference_attacks.npz holds what the reference's own attacks.zeroth_order_gradient (:17-45) and
    FGSM(order=0) (:81-102) return on the NumPy stand-in for TensorFlow (make_reference_attack_fixtures.py).  The drop-in
    functions need nothing but model.predict, so they are checked against it here on the CPU with the same model object."""
    import sys
    sys.path.insert(0, golden_dir)
    from make_reference_attack_fixtures import FixtureModel
    from deepbayes_b200 import analyzers
    g = np.load(os.path.join(golden_dir, "reference_attacks.npz"))
    L = oracle.mlp([int(v) for v in g["dims"]], hidden_act="tanh")
    model = FixtureModel(oracle, L, oracle.split_flat(L, g["W_flat"]))
    step = float(g["step"][0])
    got = analyzers.zeroth_order_gradient(model, g["x"], g["labels"], None, num_models=1, step=step)
    assert got.shape == g["zog"].shape
    np.testing.assert_allclose(got, g["zog"], rtol=0, atol=2e-5 * np.abs(g["zog"]).max() + 2e-6)
    got3 = analyzers.zeroth_order_gradient(model, g["x3"], np.array([2, 3]), None, num_models=1, step=step)
    assert got3.shape == g["zog3"].shape
    np.testing.assert_allclose(got3, g["zog3"], rtol=0, atol=2e-5 * np.abs(g["zog3"]).max() + 2e-6)
    adv = analyzers.FGSM(model, g["x"], None, float(g["fgsm_eps"][0]), direction=-1, order=0, samples=1)
    np.testing.assert_allclose(adv, g["fgsm0_adv"], rtol=0, atol=1e-6)
    # and the analytic input gradient of the same weights agrees with the reference's central differences (step 0.05, tanh net)
    W = oracle.split_flat(L, g["W_flat"])
    for i in range(g["x"].shape[0]):
        want = oracle.input_gradient(L, [W], g["x"][i:i + 1], g["labels"][i:i + 1])[0]
        np.testing.assert_allclose(g["zog"][i], want, rtol=0, atol=0.03 * np.abs(want).max() + 1e-4)


def test_uncertainty_helpers_match_a_run_of_the_reference_source(oracle, golden_dir):
    """tests/golden/reference_uncertainty.npz = the reference's own uncertainty.py (:15-71) on a model object whose predict is the
    oracle's Monte-Carlo mean over Philox-keyed posterior samples.  The drop-in helpers get the same numbers from ONE pass of
    moments (E[p], E[p^2]) instead of n stacked single-sample predictions."""
    import sys
    sys.path.insert(0, golden_dir)
    from make_reference_uncertainty_fixtures import McModel
    from deepbayes_b200.analyzers import uncertainty
    g = np.load(os.path.join(golden_dir, "reference_uncertainty.npz"))
    L = oracle.mlp([20, 16, 6])
    mean, std = oracle.synthetic_posterior(L, seed=5, sigma_scale=4.0)
    n = int(g["n"][0])
    epi, ale = uncertainty.variational_uncertainty(McModel(oracle, L, mean, std), g["x"], num_samples=n)
    np.testing.assert_allclose(epi, g["epistemic"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(ale, g["aleatoric"], rtol=0, atol=2e-6)
    assert (g["epistemic"] > 1e-4).any()                                       # the posterior is wide enough for the test to mean something
    ent = uncertainty.predictive_entropy(McModel(oracle, L, mean, std), g["x"], num_models=n)
    np.testing.assert_allclose(ent, g["entropy"], rtol=1e-5)
    lr = uncertainty.likelihood_ratio(McModel(oracle, L, mean, std), g["x"], g["x_out"], num_models=n)
    np.testing.assert_allclose(lr, g["likelihood_ratio"][0], rtol=1e-6)
