"""Seeded parity cases shared by make_golden.py (writes expected outputs with the oracle),
tests/test_golden.py (oracle vs committed outputs, CPU) and tests/test_gpu_parity.py (CUDA
path vs oracle and vs committed outputs).  Inputs are regenerated from the seeds, only the
expected outputs are stored."""
import numpy as np

from oracle import oracle as o


def case_inputs(name):
    if name == "mlp_small":      # ragged dims on purpose (not multiples of the GPU tiles)
        layers = o.mlp([37, 50, 10]); B, n = 19, 6
    elif name == "mlp_cfg1":     # BASELINE config 1 architecture, small B / n
        layers = o.mlp([784, 512, 10]); B, n = 33, 4
    elif name == "mlp_cfg2":     # BASELINE config 2 architecture, small B / n
        layers = o.mlp([784, 512, 512, 10]); B, n = 130, 3
    elif name == "mlp_tanh_reg":  # regression head: linear output, tanh / sigmoid hidden
        layers = [o.Dense(13, 24, "tanh"), o.Dense(24, 16, "sigmoid"), o.Dense(16, 3, "linear")]; B, n = 9, 5
    elif name == "cnn_small":
        layers = [o.Conv2D(3, 3, 3, 8, "relu"), o.Conv2D(3, 3, 8, 12, "relu", strides=(2, 2), padding="same"),
                  o.MaxPool2D((2, 2)), o.Flatten(), o.Dense(3 * 3 * 12, 20, "relu"), o.Dense(20, 10, "softmax")]
        B, n = 5, 3
    else:
        raise KeyError(name)
    if name == "cnn_small":
        x = o.synthetic_x((B, 14, 14, 3), seed=10)
        in_shape = (14, 14, 3)
    else:
        x = o.synthetic_x((B, layers[0].fan_in), seed=10)
        in_shape = (layers[0].fan_in,)
    mean, std = o.synthetic_posterior(layers, seed=11, sigma_scale=0.3)
    eps = o.synthetic_eps(layers, n, seed=12)
    labels = np.random.Generator(np.random.Philox(13)).integers(0, layers[-1].units, size=B).astype(np.int32)
    return dict(layers=layers, x=x, mean=mean, std=std, eps=eps, n=n, B=B, labels=labels, in_shape=in_shape)


def w_stride(layers):
    """Sampled weights are stored in full for small models, every 257th element otherwise."""
    return 1 if o.param_count(layers) < 20000 else 257


NAMES = ["mlp_small", "mlp_cfg1", "mlp_cfg2", "mlp_tanh_reg", "cnn_small"]


def expected(name):
    c = case_inputs(name)
    L, x, mean, std, eps, n = c["layers"], c["x"], c["mean"], c["std"], c["eps"], c["n"]
    out = {}
    out["predict_fp32"] = o.predict(L, mean, std, x, n, eps)
    out["predict_logits_fp32"] = o.predict(L, mean, std, x, n, eps, out="logits")
    out["predict_bf16"] = o.predict(L, mean, std, x, n, eps, bf16=True)
    out["predict_inflate2_fp32"] = o.predict(L, mean, std, x, n, eps, inflate=2.0)
    s1, s2 = o.predict_moments(L, mean, std, x, n, eps)
    out["sum1_fp32"], out["sum2_fp32"] = s1, s2
    rho = [o.inverse_softplus(s) for s in std]
    out["bbb_fp32"], w = o.bbb_forward(L, mean, rho, eps[0], x)
    st = w_stride(L)
    out["bbb_w"] = o.flatten_list(w)[::st]
    samples = [o.sample_gaussian(mean, std, e) for e in eps]
    out["sample_w"] = np.stack([o.flatten_list(s) for s in samples])[:, ::st]
    freq = o.normalise_freq(np.arange(1, n + 1))
    out["stored_expect_fp32"] = o.predict_stored(L, samples, x, list(range(n)), weights=freq)
    idx = [(3 * j + 1) % n for j in range(2 * n)]
    out["stored_mc_fp32"] = o.predict_stored(L, samples, x, idx) / np.float32(len(idx))
    flags, counts = o.count_predicate(L, mean, std, x, c["labels"], n, eps)
    out["flags"], out["counts"] = flags, counts
    return out
