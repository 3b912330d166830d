"""GPU parity at the BASELINE.json shapes: the tensor-core kernels against the CPU ORACLE (not
against another device path) on identical seeded inputs with the oracle's eps injected, through
the C ABI.

  config 2  784-512-512-10, B = 4096 (Q = 16 row-tile pairs, [s][q] work order, 74 clusters, running
            sums persisting over several chunk launches, finish4_kernel)       -> fused_mlp2_kernel
  config 3  784-1024-1024-10 stored posterior, B = 64 / B = 512                -> dense_tc_f32w_kernel /
                                                                                  dense_tc_pair_kernel
  config 4  CNN of SURVEY 8(d) on [B,32,32,3]                                   -> conv_tc_kernel + pooled epilogue
  config 5  784-512-512-10, K = 128 fixed inputs, argmax predicate              -> fused_mlp_kernel<1>
  config 1  784-512-10 at B = 10000 rows                                         -> fused_mlp2_kernel, one hidden layer
  config 5 / config 2 shapes through the widened rows in bf16 mode: expected input gradient + per-sample FGSM
            (dense_tc_bwd_kernel, FGSM epilogue) and one training step (dense_tc_wgrad_kernel)

Tolerances (BASELINE.json north_star): bf16 mode <= 2e-2 absolute on predictive probabilities
against the fp32 oracle, <= 2e-3 against the oracle that rounds the GEMM operands to bf16 (only
the accumulation order differs), argmax identical; fp32 mode <= 1e-5 relative.
Reference loop: deepbayes/posteriormodel.py:81-101; stored posteriors hmc.py:308-322;
verification loop analyzers/verifiers.py:588-634."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

BF16_ATOL = 2e-2
BF16_VS_BF16_ORACLE = 2e-3


def _mlp(dims, oracle, seed=1, sigma_scale=0.1):
    import deepbayes_b200 as db
    from deepbayes_b200.engine import Engine
    L = oracle.mlp(dims)
    layers = [db.Dense(dims[1], activation="relu", input_shape=(dims[0],))]
    for h in dims[2:-1]:
        layers.append(db.Dense(h, activation="relu"))
    layers.append(db.Dense(dims[-1], activation="softmax"))
    eng = Engine(db.Sequential(layers))
    mean, std = oracle.synthetic_posterior(L, seed=seed, sigma_scale=sigma_scale)
    return L, eng, mean, std


def _check_probs(got, want_bf, want_32):
    assert np.isfinite(got).all()
    assert np.abs(got - want_32).max() <= BF16_ATOL
    np.testing.assert_allclose(got, want_bf, rtol=0, atol=BF16_VS_BF16_ORACLE)
    assert (got.argmax(-1) == want_bf.argmax(-1)).all(), "argmax differs from the bf16 oracle"
    margin = np.sort(want_32, axis=-1)
    clear = (margin[:, -1] - margin[:, -2]) > 2 * BF16_ATOL
    assert (got.argmax(-1) == want_32.argmax(-1))[clear].all()


@pytest.mark.parametrize("fused_ns", [None, 4, 3])
def test_config2_full_batch_fused_kernel_vs_oracle(oracle, dbx_opt, fused_ns):
    """BASELINE config 2 at its real batch: B = 4096 rows = 16 row-tile pairs per sample.  fused_ns = 4 / 3 cut the 8 samples
    into 2 / 3 launches, so the per-(cluster, q) running sums persist across launches and the last chunk is ragged."""
    dims, B, n = [784, 512, 512, 10], 4096, 8
    L, eng, mean, std = _mlp(dims, oracle)
    eng.set_gaussian(mean, std)
    x = oracle.synthetic_x((B, dims[0]), seed=0)
    eps = oracle.synthetic_eps(L, n, seed=2)
    eps_flat = np.stack([oracle.flatten_list(e) for e in eps])
    if fused_ns is not None:
        dbx_opt("FUSED_NS", fused_ns)
    s1, s2 = eng.predict_sums(x, n, eps=eps_flat, mode="bf16", second_moment=True)
    assert eng.last_path() == "fused_tcgen05"
    got = (s1 / float(n)).cpu().numpy()
    w1, w2 = oracle.predict_moments(L, mean, std, x, n, eps, bf16=True)
    want_32 = oracle.predict(L, mean, std, x, n, eps)
    _check_probs(got, w1 / np.float32(n), want_32)
    np.testing.assert_allclose((s2 / float(n)).cpu().numpy(), w2 / np.float32(n), rtol=0, atol=BF16_VS_BF16_ORACLE)
    # mean logits (predict_logits, posteriormodel.py:114-139) through the same kernel
    lg, _ = eng.predict_sums(x, n, eps=eps_flat, mode="bf16", out_kind="logits")
    want_lg = oracle.predict(L, mean, std, x, n, eps, out="logits", bf16=True)
    np.testing.assert_allclose((lg / float(n)).cpu().numpy(), want_lg, rtol=0, atol=5e-3)


def test_config2_full_batch_fp32_mode_vs_oracle(oracle):
    dims, B, n = [784, 512, 512, 10], 4096, 3
    L, eng, mean, std = _mlp(dims, oracle)
    eng.set_gaussian(mean, std)
    x = oracle.synthetic_x((B, dims[0]), seed=0)
    eps = oracle.synthetic_eps(L, n, seed=2)
    eps_flat = np.stack([oracle.flatten_list(e) for e in eps])
    s1, _ = eng.predict_sums(x, n, eps=eps_flat, mode="fp32")
    want = oracle.predict(L, mean, std, x, n, eps)
    np.testing.assert_allclose((s1 / float(n)).cpu().numpy(), want, rtol=1e-5, atol=2e-7)


def test_config1_full_batch_vs_oracle(oracle):
    """BASELINE config 1 (784-512-10) at the MNIST-test-shaped batch of SURVEY 8(d): B = 10,000 (79 row tiles: the last pair is
    half empty and the last tile ragged)."""
    dims, B, n = [784, 512, 10], 10000, 5
    L, eng, mean, std = _mlp(dims, oracle)
    eng.set_gaussian(mean, std)
    x = oracle.synthetic_x((B, dims[0]), seed=0)
    eps = oracle.synthetic_eps(L, n, seed=2)
    eps_flat = np.stack([oracle.flatten_list(e) for e in eps])
    s1, _ = eng.predict_sums(x, n, eps=eps_flat, mode="bf16")
    assert eng.last_path() == "fused_tcgen05"
    _check_probs((s1 / float(n)).cpu().numpy(), oracle.predict(L, mean, std, x, n, eps, bf16=True),
                 oracle.predict(L, mean, std, x, n, eps))


@pytest.mark.parametrize("B", [64, 512])
def test_config3_stored_posterior_vs_oracle(oracle, B):
    """BASELINE config 3 architecture (784-1024-1024-10), stored fp32 samples resident in HBM: B = 64 runs
    dense_tc_f32w_kernel (weights by TMA straight from the fp32 slab), B = 512 the conversion pass + dense_tc_pair_kernel."""
    dims, S = [784, 1024, 1024, 10], 8
    L, eng, mean, std = _mlp(dims, oracle, sigma_scale=0.3)
    eps = oracle.synthetic_eps(L, S, seed=2)
    samples = [oracle.sample_gaussian(mean, std, e) for e in eps]
    slab = np.stack([oracle.flatten_list(s) for s in samples])
    eng.set_samples(slab)
    x = oracle.synthetic_x((B, dims[0]), seed=0)
    g = np.random.Generator(np.random.Philox(5))
    idx = g.integers(0, S, size=12).astype(np.int32)          # the n draws of posteriormodel.py:77, with replacement
    s1, _ = eng.predict_stored_sums(x, len(idx), idx=idx, mode="bf16")
    assert eng.last_path() == "generic_tc"
    want_bf = oracle.predict_stored(L, samples, x, list(idx), bf16=True) / np.float32(len(idx))
    want_32 = oracle.predict_stored(L, samples, x, list(idx)) / np.float32(len(idx))
    _check_probs((s1 / float(len(idx))).cpu().numpy(), want_bf, want_32)
    # deterministic expectation sum_i freq_i f(W_i) (probverification.py:619-650), contiguous rows
    freq = oracle.normalise_freq(np.arange(1, S + 1))
    e1, _ = eng.predict_stored_sums(x, S, j0=0, wts=freq, mode="bf16")
    want_e = oracle.predict_stored(L, samples, x, list(range(S)), weights=freq, bf16=True)
    np.testing.assert_allclose(e1.cpu().numpy(), want_e, rtol=0, atol=BF16_VS_BF16_ORACLE)
    # fp32 mode: the reference's arithmetic
    f1, _ = eng.predict_stored_sums(x[:16], S, j0=0, wts=freq, mode="fp32")
    np.testing.assert_allclose(f1.cpu().numpy(), oracle.predict_stored(L, samples, x[:16], list(range(S)), weights=freq),
                               rtol=1e-5, atol=2e-7)


def test_config4_cnn_vs_oracle(oracle):
    """BASELINE config 4: Conv2D(32,3x3) -> Conv2D(64,3x3) -> MaxPool(2) -> Flatten -> Dense(128) -> Dense(10) on [B,32,32,3]
    (SURVEY 8(d)): first conv through the shared im2col + GEMM, second through conv_tc_kernel with the pooled epilogue."""
    import deepbayes_b200 as db
    from deepbayes_b200.engine import Engine
    L = oracle.cnn_cfg4()
    m = db.Sequential([db.Conv2D(32, 3, activation="relu", input_shape=(32, 32, 3)), db.Conv2D(64, 3, activation="relu"),
                       db.MaxPooling2D(2), db.Flatten(), db.Dense(128, activation="relu"), db.Dense(10, activation="softmax")])
    eng = Engine(m)
    assert eng.P == oracle.param_count(L) == 1626442
    mean, std = oracle.synthetic_posterior(L, seed=1)
    eng.set_gaussian(mean, std)
    B, n = 4, 2
    x = oracle.synthetic_x((B, 32, 32, 3), seed=0)
    eps = oracle.synthetic_eps(L, n, seed=2)
    eps_flat = np.stack([oracle.flatten_list(e) for e in eps])
    s1, _ = eng.predict_sums(x, n, eps=eps_flat, mode="bf16")
    assert eng.last_path() == "generic_tc"
    _check_probs((s1 / float(n)).cpu().numpy(), oracle.predict(L, mean, std, x, n, eps, bf16=True),
                 oracle.predict(L, mean, std, x, n, eps))
    f1, _ = eng.predict_sums(x, n, eps=eps_flat, mode="fp32")
    np.testing.assert_allclose((f1 / float(n)).cpu().numpy(), oracle.predict(L, mean, std, x, n, eps), rtol=1e-5, atol=2e-7)


def test_config5_predicate_flags_vs_oracle(oracle):
    """BASELINE config 5: K = 128 fixed (pre-perturbed) inputs through config 2's MLP, per-sample predicate argmax == label
    (verifiers.py:628-634 with the tutorial predicate).  Flags must equal the bf16 oracle's wherever its top-2 logits are not
    within accumulation-order noise of each other."""
    dims, K, n = [784, 512, 512, 10], 128, 16
    L, eng, mean, std = _mlp(dims, oracle)
    eng.set_gaussian(mean, std)
    g = np.random.Generator(np.random.Philox(3))
    x = np.clip(oracle.synthetic_x((K, dims[0]), seed=0) + g.uniform(-0.1, 0.1, size=(K, dims[0])).astype(np.float32), 0, 1)
    eps = oracle.synthetic_eps(L, n, seed=2)
    eps_flat = np.stack([oracle.flatten_list(e) for e in eps])
    # labels = the deterministic prediction, so that the predicate is not trivially false
    labels = oracle.forward(L, mean, x).argmax(-1).astype(np.int32)
    counts, flags = eng.count_predicate(x, labels, n, eps=eps_flat, mode="bf16", return_flags=True)
    assert eng.last_path() == "fused_tcgen05"
    flags, counts = flags.cpu().numpy(), counts.cpu().numpy()
    np.testing.assert_array_equal(counts, flags.sum(0))
    want_flags = np.zeros((n, K), np.uint8)
    clear = np.zeros((n, K), bool)
    for s in range(n):
        w = oracle.sample_gaussian(mean, std, eps[s])
        z = oracle.forward(L, w, x, np.float32, "logits", True)
        want_flags[s] = (z.argmax(-1) == labels)
        top = np.sort(z, axis=-1)
        clear[s] = (top[:, -1] - top[:, -2]) > 5e-3
    assert clear.mean() > 0.9
    np.testing.assert_array_equal(flags[clear], want_flags[clear])
    assert (flags != want_flags).mean() <= 0.01
    assert 0 < counts.sum() <= n * K


def test_rank_shards_equal_one_rank(oracle):
    """SURVEY 8(e) on one GPU: the n samples of one predict split into G contiguous global-id ranges (what rank g of G
    evaluates) and summed equal the single-rank result, for G = 2, 4, 8 (bench.py repeats this across real ranks under
    torchrun and prints "shard_check")."""
    from deepbayes_b200 import distributed as dd
    dims, B, n = [784, 512, 512, 10], 1024, 64
    L, eng, mean, std = _mlp(dims, oracle)
    eng.set_gaussian(mean, std)
    x = oracle.synthetic_x((B, dims[0]), seed=0)
    one1, one2 = eng.predict_sums(x, n, seed=11, s0=1000, mode="bf16", second_moment=True)
    one1, one2 = one1.cpu().numpy(), one2.cpu().numpy()
    for G in (2, 4, 8):
        t1 = np.zeros_like(one1); t2 = np.zeros_like(one2)
        for g in range(G):
            off, cnt = dd.shard_range(n, g, G)
            a, b = eng.predict_sums(x, cnt, seed=11, s0=1000 + off, mode="bf16", second_moment=True)
            t1 += a.cpu().numpy(); t2 += b.cpu().numpy()
        np.testing.assert_allclose(t1 / n, one1 / n, rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(t2 / n, one2 / n, rtol=1e-5, atol=1e-6)


def test_stream_mode_equals_chunked_launches(oracle, dbx_opt):
    """Stream mode (one fused launch per chunk, sampler running concurrently behind per-sub-chunk ready flags) against the
    event-ordered path (FUSED_STREAM=0: sampler of a chunk finished before its fused launch), the output-layer MMAs slipped
    into the next chunk's stream or issued after it, and the two work orders: the same sums up to the order the per-cluster
    slices are added in."""
    dims, B, n = [784, 512, 512, 10], 2048, 96      # 8 row-tile pairs per sample: stream mode is the default from 6 up
    L, eng, mean, std = _mlp(dims, oracle)
    eng.set_gaussian(mean, std)
    x = oracle.synthetic_x((B, dims[0]), seed=0)
    ref1, ref2 = eng.predict_sums(x, n, seed=5, s0=40, mode="bf16", second_moment=True)
    assert eng.last_path() == "fused_tcgen05"
    ref1, ref2 = ref1.cpu().numpy() / n, ref2.cpu().numpy() / n
    for opts in ({"FUSED_STREAM": 0}, {"FUSED_L3AT": 0}, {"FUSED_L3AT": 24}, {"FUSED_ORDER": 0}, {"FUSED_PRIO": 0, "FUSED_STREAM": 0}):
        for k, v in opts.items():
            dbx_opt(k, v)
        a1, a2 = eng.predict_sums(x, n, seed=5, s0=40, mode="bf16", second_moment=True)
        for k in opts:
            dbx_opt(k, None)
        np.testing.assert_allclose(a1.cpu().numpy() / n, ref1, rtol=2e-6, atol=2e-7, err_msg=str(opts))
        np.testing.assert_allclose(a2.cpu().numpy() / n, ref2, rtol=2e-6, atol=2e-7, err_msg=str(opts))
    # more than two chunks of 1024 samples: both weight buffers, their flags and the cross-chunk running sums in play
    big1, big2 = eng.predict_sums(x, 2100, seed=6, s0=7, mode="bf16", second_moment=True)
    dbx_opt("FUSED_STREAM", 0)
    ref_b1, ref_b2 = eng.predict_sums(x, 2100, seed=6, s0=7, mode="bf16", second_moment=True)
    dbx_opt("FUSED_STREAM", None)
    np.testing.assert_allclose(big1.cpu().numpy() / 2100, ref_b1.cpu().numpy() / 2100, rtol=2e-6, atol=2e-7)
    np.testing.assert_allclose(big2.cpu().numpy() / 2100, ref_b2.cpu().numpy() / 2100, rtol=2e-6, atol=2e-7)
    # back-to-back calls of different sizes in the default mode (buffers alternate from call to call)
    for n2 in (8, 200, 17, 1500):
        d1, _ = eng.predict_sums(x, n2, seed=9, s0=3, mode="bf16")
        dbx_opt("FUSED_STREAM", 0)
        e1, _ = eng.predict_sums(x, n2, seed=9, s0=3, mode="bf16")
        dbx_opt("FUSED_STREAM", None)
        np.testing.assert_allclose(d1.cpu().numpy() / n2, e1.cpu().numpy() / n2, rtol=2e-6, atol=2e-7)
    # back-to-back calls reuse the weight buffers and their ready flags: results must not depend on what ran before
    for n2 in (8, 200, 17):
        dbx_opt("FUSED_STREAM", 2)      # forced also for a batch this small (2 row-tile pairs)
        b1, _ = eng.predict_sums(x[:300], n2, seed=9, s0=3, mode="bf16")
        dbx_opt("FUSED_STREAM", 0)
        c1, _ = eng.predict_sums(x[:300], n2, seed=9, s0=3, mode="bf16")
        dbx_opt("FUSED_STREAM", None)
        np.testing.assert_allclose(b1.cpu().numpy() / n2, c1.cpu().numpy() / n2, rtol=2e-6, atol=2e-7)


def test_f32w_wide_tiles_equal_narrow(oracle, dbx_opt):
    """dense_tc_f32w_kernel with 128-column tiles / 64-row A boxes against its 64-column form: the same MMAs in the same k
    order, so the sums are bit-identical."""
    dims, S = [784, 1024, 1024, 10], 6
    L, eng, mean, std = _mlp(dims, oracle, sigma_scale=0.3)
    eng.set_gaussian(mean, std)
    eng.set_samples(eng.sample(S, seed=3, s0=0))
    for B in (64, 100, 33):
        x = oracle.synthetic_x((B, dims[0]), seed=B)
        a1, a2 = eng.predict_stored_sums(x, S, j0=0, mode="bf16", second_moment=True)
        assert eng.last_path() == "generic_tc"
        dbx_opt("NO_F32W_WIDE", 1)
        b1, b2 = eng.predict_stored_sums(x, S, j0=0, mode="bf16", second_moment=True)
        dbx_opt("NO_F32W_WIDE", None)
        np.testing.assert_array_equal(a1.cpu().numpy(), b1.cpu().numpy())
        np.testing.assert_array_equal(a2.cpu().numpy(), b2.cpu().numpy())


def test_peer_window_allreduce_one_rank():
    """dbx_comm_* on a world of one rank: the kernel's copy-in / flag / rank-ordered sum path with a single window (the
    N-rank exchange is exercised by bench.py under torchrun, which asserts it against a one-rank recomputation)."""
    import torch
    from deepbayes_b200 import distributed as dd
    c = dd.PeerComm(0, 1, torch.cuda.current_device(), 1 << 17)
    try:
        c.connect([c.handle()])
        a = torch.arange(4096 * 10, dtype=torch.float32, device="cuda") * 0.5
        b = torch.full((4096 * 10,), 3.0, dtype=torch.float32, device="cuda")
        wa, wb = a.clone(), b.clone()
        for _ in range(5):      # epochs alternate the two slots of the window
            c.allreduce(a, b)
        torch.cuda.synchronize()
        assert torch.equal(a, wa) and torch.equal(b, wb) and not c.timed_out()
        c.allreduce(a)
        torch.cuda.synchronize()
        assert torch.equal(a, wa)
        from deepbayes_b200._lib import DbxError
        with pytest.raises(DbxError):
            c.allreduce(torch.zeros(1 << 18, dtype=torch.float32, device="cuda"))   # larger than the window
    finally:
        c.close()


@pytest.mark.parametrize("dims", [[784, 384, 384, 10], [784, 128, 10], [784, 256, 256, 7], [200, 512, 512, 16]])
def test_other_fused_shapes_in_stream_mode_vs_oracle(oracle, dims):
    """The pair kernel's other geometries (hidden width 128 / 256 / 384: a 128-column second layer-1 chunk, one or three
    layer-2 k-pieces, TMA boxes that reach past the kernel's columns; one hidden layer; K0 not a multiple of 128) at a batch
    large enough for stream mode (7 row-tile pairs, the last one ragged), against the oracle."""
    B, n = 1700, 5
    L, eng, mean, std = _mlp(dims, oracle)
    eng.set_gaussian(mean, std)
    x = oracle.synthetic_x((B, dims[0]), seed=0)
    eps = oracle.synthetic_eps(L, n, seed=2)
    eps_flat = np.stack([oracle.flatten_list(e) for e in eps])
    s1, s2 = eng.predict_sums(x, n, eps=eps_flat, mode="bf16", second_moment=True)
    assert eng.last_path() == "fused_tcgen05"
    w1, w2 = oracle.predict_moments(L, mean, std, x, n, eps, bf16=True)
    _check_probs((s1 / float(n)).cpu().numpy(), w1 / np.float32(n), oracle.predict(L, mean, std, x, n, eps))
    np.testing.assert_allclose((s2 / float(n)).cpu().numpy(), w2 / np.float32(n), rtol=0, atol=BF16_VS_BF16_ORACLE)


def test_config5_shape_input_gradient_and_fgsm_on_tensor_cores_vs_oracle(oracle):
    """Config 5's shape (784-512-512-10, K = 128 inputs) through the gradient loops in bf16 mode: the expected input gradient of
    attacks.py:49-75 (2 iterations x 8 posterior samples) and the per-sample FGSM walk of verifiers.py:622-634 (12 samples),
    forward / backward on tcgen05 (dense_tc_persist_kernel, dense_tc_bwd_kernel with the FGSM step in its epilogue), against the
    oracle that rounds the GEMM operands the same way and against the exact (fp64) gradient."""
    dims, K = [784, 512, 512, 10], 128
    L, eng, mean, std = _mlp(dims, oracle)
    eng.set_gaussian(mean, std)
    x = oracle.synthetic_x((K, dims[0]), seed=0)
    labels = oracle.forward(L, mean, x).argmax(-1).astype(np.int32)
    n_outer, n_inner = 2, 8
    eps = oracle.synthetic_eps(L, n_outer * n_inner, seed=2)
    eps_flat = np.stack([oracle.flatten_list(e) for e in eps])
    got = eng.input_gradient(x, labels, n_outer, n_inner, noise=eps_flat, mode="bf16").cpu().numpy()
    assert eng.last_path() == "generic_tc"
    want16, want = np.zeros((K, dims[0])), np.zeros((K, dims[0]))
    for it in range(n_outer):
        ws = [oracle.sample_gaussian(mean, std, eps[it * n_inner + s], order="f32") for s in range(n_inner)]
        want16 += oracle.input_gradient(L, ws, x, labels, bf16=True)
        want += oracle.input_gradient(L, ws, x, labels)
    scale = np.abs(want).max()
    row_err = np.abs(got - want16).max(1)
    assert np.median(row_err) <= 1e-3 * scale and row_err.max() <= 0.1 * scale, (np.median(row_err) / scale, row_err.max() / scale)
    assert np.linalg.norm(got - want) <= 0.1 * np.linalg.norm(want)
    # the per-sample FGSM loop: flags against the oracle walk wherever neither the gradient signs nor the top-2 logits are near a tie
    n, e = 12, 0.05
    counts, flags = eng.count_predicate_fgsm(x, labels, labels, e, n, noise=eps_flat[:n], lower=0.0, upper=1.0, return_flags=True, mode="bf16")
    assert eng.last_path() == "generic_tc"
    flags = flags.cpu().numpy().astype(bool)
    np.testing.assert_array_equal(counts.cpu().numpy(), flags.sum(0))
    want_flags, clear = np.zeros((n, K), bool), np.zeros((n, K), bool)
    for s in range(n):
        w = oracle.sample_gaussian(mean, std, eps[s], order="f32")
        gr = oracle.input_gradient(L, [w], x, labels, bf16=True)
        adv = oracle.fgsm(oracle.bf16_round(x), gr, e, 0.0, 1.0)
        z = oracle.forward(L, w, adv, np.float32, "logits", True)
        want_flags[s] = z.argmax(-1) == labels
        top = np.sort(z, axis=-1)
        near_zero = (np.abs(gr) < 2e-2 * np.abs(gr).max(1, keepdims=True)).mean(1)
        clear[s] = ((top[:, -1] - top[:, -2]) > 0.05 * np.abs(z).max()) & (near_zero < 0.5)
    assert clear.mean() > 0.5 and (flags == want_flags)[clear].mean() > 0.97


def test_config2_shape_training_step_on_tensor_cores_vs_oracle(oracle):
    """One Bayes-by-Backprop step (bayesbybackprop.py:46-170) of config 2's MLP at mini-batch 512 in bf16 training mode:
    dense_tc_kernel / dense_tc_bwd_kernel / dense_tc_wgrad_kernel (dW tiles with a ragged last row block: 784 = 6 x 128 + 16)
    against the oracle with the same operand roundings and against the exact step."""
    import deepbayes_b200 as db
    dims, B = [784, 512, 512, 10], 512
    L = oracle.mlp(dims)
    layers = [db.Dense(512, activation="relu", input_shape=(784,)), db.Dense(512, activation="relu"), db.Dense(10, activation="softmax")]
    opt = db.optimizers.BayesByBackprop().compile(db.Sequential(layers), None, batch_size=B, inflate_prior=1.0, kl_weight=1.0, train_mode="bf16")
    mean, std = oracle.synthetic_posterior(L, seed=1, sigma_scale=0.1)
    opt.posterior_mean = [m.copy() for m in mean]
    opt.posterior_var = [oracle.inverse_softplus(s) for s in std]
    rho0 = [r.copy() for r in opt.posterior_var]
    x = oracle.synthetic_x((B, 784), seed=0)
    lab = np.random.Generator(np.random.Philox(3)).integers(0, 10, size=B)
    noise = oracle.synthetic_eps(L, 1, seed=2)[0]
    nm, nr = opt.step(x, lab, 0.1, noise=noise)
    assert opt.engine().last_path() == "generic_tc"
    wm16, wr16, loss16, _, pred16, _ = oracle.bbb_step(L, mean, rho0, opt.prior_mean, opt.prior_var, x, lab, noise, 0.1, dtype=np.float32, bf16=True)
    wm, wr, loss, klc, pred, W = oracle.bbb_step(L, mean, rho0, opt.prior_mean, opt.prior_var, x, lab, noise, 0.1)
    for i in range(len(nm)):
        d16, d64, dg = [(a - mean[i]).ravel().astype(np.float64) for a in (wm16[i], wm[i], nm[i])]
        # (what is left against the same-rounding oracle: fp32 summation order moving stored activations / upstream gradients across bf16 ties)
        assert np.linalg.norm(dg - d16) <= 5e-2 * np.linalg.norm(d16) + 1e-9, (i, np.linalg.norm(dg - d16) / np.linalg.norm(d16))
        assert np.linalg.norm(dg - d64) <= 0.1 * np.linalg.norm(d64) + 1e-9, (i, np.linalg.norm(dg - d64) / np.linalg.norm(d64))
        # rho = -5.6 here and one step moves it by about one fp32 ulp (4.8e-7): the fp32 step (the reference's arithmetic too) is
        # compared in ulps, not relative to the update
        ulp = np.spacing(np.abs(rho0[i]).max().astype(np.float32))
        assert (np.abs(nr[i] - wr16[i]) <= 2 * ulp).mean() > 0.999 and np.abs(nr[i] - wr16[i]).max() <= 16 * ulp
        assert np.abs(nr[i].astype(np.float64) - wr[i]).max() <= 16 * ulp
        np.testing.assert_allclose(opt.model.get_weights()[i], W[i], rtol=1e-6, atol=1e-7)
    probs = opt._last_probs.cpu().numpy()
    np.testing.assert_allclose(probs, pred16, rtol=0, atol=BF16_VS_BF16_ORACLE)
    assert np.abs(probs - pred).max() <= BF16_ATOL
    top = np.sort(pred, axis=1)
    clear = (top[:, -1] - top[:, -2]) > 2e-3          # (an untrained model: many rows are near-ties)
    assert (probs.argmax(1) == pred.argmax(1))[clear].all()
    np.testing.assert_allclose(opt._last_loss.cpu().numpy()[0], loss16, rtol=2e-3)
