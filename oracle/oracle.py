"""CPU oracle for the deepbayes posterior-predictive hot path.

*** TEST INFRASTRUCTURE ONLY ***
Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import this module.  The product (``deepbayes_b200``) never does: it
calls the sm_100a CUDA library through the C ABI and fails loudly when that library is missing.

What this is: a loop-faithful NumPy restatement of the reference's algorithm for the path
named by BASELINE.json ``north_star`` (Monte-Carlo posterior predictive + Bayes-by-Backprop
reparameterised forward).  Every function cites the reference lines it follows
(paths relative to /root/reference).

Parity status: the reference ships NO tests and NO known-answer vectors for this path and it
cannot be imported here as-is (TensorFlow / Keras / TFP are not installed, there is no
network).  The oracle is therefore pinned two ways, both weaker than a real golden vector:
  (1) ``tests/golden/make_reference_fixtures.py`` executes the reference's OWN source files
      (posteriormodel.py, optimizers/optimizer.py, optimizers/bayesbybackprop.py, loaded
      unmodified from /root/reference) on top of a NumPy stand-in for the handful of TF leaf
      ops they call; the committed outputs pin the reference's control flow (RNG consumption
      order, var-used-as-std, fp32 running sum, /float(n), softplus(rho) ...).
  (2) the two saved posteriors under Tutorials/PosteriorModels pin real inputs (mu, sigma,
      architecture).
  (3) the widened rows: make_reference_ibp_fixtures.py runs analyzers/verifiers.py (IBP,
      chernoff_bound_verification), make_reference_conv_ibp_fixtures.py runs its propagate_conv2d (with a NumPy
      VALID / stride-1 loop for tf.nn.convolution) and make_reference_bbb_fixtures.py runs optimizers/losses.py
      (log_q, KL_Loss) the same way; the GRADIENTS of bbb_step / input_gradient replace
      tf.GradientTape, which cannot run on the stand-in, and are checked against finite
      differences of those pinned losses -- parity unpinned by a reference run for them.
  (4) make_reference_attack_fixtures.py runs attacks.zeroth_order_gradient / FGSM(order=0) (they need only tf.eye,
      tf.reshape and model.predict): the analytic input_gradient below agrees with those reference-run central
      differences; make_reference_uncertainty_fixtures.py runs analyzers/uncertainty.py;
      make_reference_massart_fixtures.py runs verifiers.massart_bound_check / okamoto_bound / absolute_massart_halting
      and records every iteration's interval bounds, which the drop-in loop replays (tests/test_host.py).
The third-party arithmetic itself (TensorFlow's Eigen matmul rounding, tf.random.normal's
bit stream) is NOT pinned by anything the reference holds: **parity unpinned** for those,
as DESIGN.md says.
"""
from __future__ import annotations

import math
from typing import Callable, List, Optional, Sequence

import numpy as np

# --------------------------------------------------------------------------------------
# Layer descriptors (mirror of the Keras Sequential pieces the reference touches, a12)
# --------------------------------------------------------------------------------------


class Dense:
    """Keras Dense: y = act(x @ W + b) on the last axis (posteriormodel.py:98 via model())."""

    def __init__(self, fan_in: int, units: int, activation: str = "linear"):
        self.kind = "dense"
        self.fan_in, self.units, self.activation = int(fan_in), int(units), activation

    def weight_shapes(self):
        return [(self.fan_in, self.units), (self.units,)]


class Conv2D:
    """Keras Conv2D, NHWC activations / HWIO kernel (verifiers.py:11-21 uses
    tf.nn.convolution, stride 1 VALID; strides / SAME are restated for completeness)."""

    def __init__(self, kh, kw, cin, cout, activation="linear", strides=(1, 1), padding="valid"):
        self.kind = "conv2d"
        self.kh, self.kw, self.cin, self.cout = int(kh), int(kw), int(cin), int(cout)
        self.activation = activation
        self.strides = (int(strides[0]), int(strides[1]))
        self.padding = padding.lower()

    def weight_shapes(self):
        return [(self.kh, self.kw, self.cin, self.cout), (self.cout,)]


class MaxPool2D:
    def __init__(self, pool=(2, 2), strides=None):
        self.kind = "maxpool2d"
        self.pool = (int(pool[0]), int(pool[1]))
        self.strides = self.pool if strides is None else (int(strides[0]), int(strides[1]))

    def weight_shapes(self):
        return []


class Flatten:
    def __init__(self):
        self.kind = "flatten"

    def weight_shapes(self):
        return []


def mlp(dims: Sequence[int], hidden_act="relu", out_act="softmax") -> list:
    """784-512-512-10 style MLP as used by BASELINE.json configs 1/2/3/5."""
    layers = []
    for i in range(len(dims) - 1):
        act = out_act if i == len(dims) - 2 else hidden_act
        layers.append(Dense(dims[i], dims[i + 1], act))
    return layers


def cnn_cfg4() -> list:
    """BASELINE config 4 architecture as fixed in SURVEY.md section 8(d)."""
    return [
        Conv2D(3, 3, 3, 32, "relu"),
        Conv2D(3, 3, 32, 64, "relu"),
        MaxPool2D((2, 2)),
        Flatten(),
        Dense(12544, 128, "relu"),
        Dense(128, 10, "softmax"),
    ]


def weight_shapes(layers) -> list:
    out = []
    for l in layers:
        out.extend(l.weight_shapes())
    return out


def param_count(layers) -> int:
    return int(sum(int(np.prod(s)) for s in weight_shapes(layers)))


def flops_per_forward(layers, input_shape=None) -> int:
    """Algorithmic FLOPs of one (sample x input row) forward, SURVEY.md 8(d)."""
    f = 0
    shape = tuple(input_shape) if input_shape is not None else None
    for l in layers:
        if l.kind == "dense":
            f += 2 * l.fan_in * l.units
            shape = (l.units,)
        elif l.kind == "conv2d":
            ho, wo = _conv_out_hw(shape[0], shape[1], l)
            f += 2 * l.kh * l.kw * l.cin * l.cout * ho * wo
            shape = (ho, wo, l.cout)
        elif l.kind == "maxpool2d":
            shape = ((shape[0] - l.pool[0]) // l.strides[0] + 1, (shape[1] - l.pool[1]) // l.strides[1] + 1, shape[2])
        elif l.kind == "flatten":
            shape = (int(np.prod(shape)),)
    return int(f)


# --------------------------------------------------------------------------------------
# Leaf arithmetic (third-party ops restated: a12, a13)
# --------------------------------------------------------------------------------------

_SOFTPLUS_THR = np.float32(np.log(np.finfo(np.float32).eps) + 2.0)  # ~ -13.9424


def softplus(x, dtype=np.float32):
    """tf.math.softplus as called at bayesbybackprop.py:22-23,54,180,189.
    TF's CPU functor: x if x > -thr; exp(x) if x < thr; else log1p(exp(x)); thr = ln(eps)+2."""
    x = np.asarray(x, dtype=dtype)
    thr = dtype(_SOFTPLUS_THR)
    with np.errstate(over="ignore"):
        ex = np.exp(x)
        mid = np.log1p(ex)
    return np.where(x > -thr, x, np.where(x < thr, ex, mid)).astype(dtype)


def inverse_softplus(std, dtype=np.float32):
    """rho = log(exp(std) - 1), bayesbybackprop.py:39-40 and optimizer.py:286."""
    std = np.asarray(std, dtype=dtype)
    return np.log(np.exp(std) - dtype(1)).astype(dtype)


def _act(name: str, z: np.ndarray) -> np.ndarray:
    if name in ("linear", None):
        return z
    if name == "relu":
        return np.maximum(z, z.dtype.type(0))
    if name == "tanh":
        return np.tanh(z)
    if name == "sigmoid":
        return (1 / (1 + np.exp(-z))).astype(z.dtype)
    if name == "softmax":
        m = z.max(axis=-1, keepdims=True)
        e = np.exp(z - m)
        return (e / e.sum(axis=-1, keepdims=True)).astype(z.dtype)
    raise ValueError("unknown activation %r" % (name,))


def _conv_out_hw(h, w, l):
    if l.padding == "same":
        return -(-h // l.strides[0]), -(-w // l.strides[1])
    return (h - l.kh) // l.strides[0] + 1, (w - l.kw) // l.strides[1] + 1


def _conv2d(x, k, l):
    """NHWC x HWIO -> NHWC, via im2col + matmul in x.dtype."""
    n, h, w, c = x.shape
    ho, wo = _conv_out_hw(h, w, l)
    sh, sw = l.strides
    if l.padding == "same":
        ph = max((ho - 1) * sh + l.kh - h, 0)
        pw = max((wo - 1) * sw + l.kw - w, 0)
        x = np.pad(x, ((0, 0), (ph // 2, ph - ph // 2), (pw // 2, pw - pw // 2), (0, 0)))
    cols = np.empty((n, ho, wo, l.kh, l.kw, c), dtype=x.dtype)
    for i in range(l.kh):
        for j in range(l.kw):
            cols[:, :, :, i, j, :] = x[:, i:i + sh * ho:sh, j:j + sw * wo:sw, :]
    y = cols.reshape(n * ho * wo, l.kh * l.kw * c) @ k.reshape(l.kh * l.kw * c, l.cout)
    return y.reshape(n, ho, wo, l.cout)


def _maxpool(x, l):
    n, h, w, c = x.shape
    ho = (h - l.pool[0]) // l.strides[0] + 1
    wo = (w - l.pool[1]) // l.strides[1] + 1
    out = np.full((n, ho, wo, c), -np.inf, dtype=x.dtype)
    for i in range(l.pool[0]):
        for j in range(l.pool[1]):
            out = np.maximum(out, x[:, i:i + l.strides[0] * ho:l.strides[0], j:j + l.strides[1] * wo:l.strides[1], :])
    return out


def bf16_round(a: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even fp32 -> bf16 -> fp32 (what cvt.rn.bf16.f32 does)."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    u = a.view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    out = r.astype(np.uint32).view(np.float32)
    return np.where(np.isnan(a), a, out).reshape(a.shape)


def forward(layers, weights, x, dtype=np.float32, out="probs", bf16=False):
    """One forward pass of the Keras Sequential with explicit weights (a12; call sites
    posteriormodel.py:93,98,100,113; bayesbybackprop.py:65).  ``out='logits'`` stops before
    the last activation (posteriormodel.py:128-134).  ``bf16=True`` emulates the GPU's bf16
    mode: GEMM operands (weights and layer inputs) rounded to bf16, fp32 accumulation, fp32
    bias / activation / softmax."""
    h = np.asarray(x).astype(dtype)
    wi = 0
    nl = len(layers)
    last_weighted = max(i for i, l in enumerate(layers) if l.kind in ("dense", "conv2d"))
    for i, l in enumerate(layers):
        if l.kind == "dense":
            W = np.asarray(weights[wi]).astype(dtype)
            b = np.asarray(weights[wi + 1]).astype(dtype)
            wi += 2
            if bf16:
                z = bf16_round(h) @ bf16_round(W) + b
            else:
                z = h @ W + b
        elif l.kind == "conv2d":
            W = np.asarray(weights[wi]).astype(dtype)
            b = np.asarray(weights[wi + 1]).astype(dtype)
            wi += 2
            if bf16:
                z = _conv2d(bf16_round(h), bf16_round(W), l) + b
            else:
                z = _conv2d(h, W, l) + b
        elif l.kind == "maxpool2d":
            h = _maxpool(h, l)
            continue
        elif l.kind == "flatten":
            h = h.reshape(h.shape[0], -1)
            continue
        else:
            raise ValueError(l.kind)
        if i == last_weighted and out == "logits":
            return z
        h = _act(l.activation, z)
    return h


# --------------------------------------------------------------------------------------
# Sampling (a2, a6, a8, a9)
# --------------------------------------------------------------------------------------


def sample_gaussian(mean, std, eps, inflate=1.0, order="f64"):
    """W = mean + (inflate*std) * eps with eps injected.
    posteriormodel.py:74-75 / optimizer.py:200-201: ``np.random.normal(loc=mean_i,
    scale=inflate*var_i)`` -- the stored ``var`` array is used AS A STANDARD DEVIATION and the
    draw is formed in float64 then cast to fp32 by set_weights (order='f64').
    bayesbybackprop.py:52-56 forms it entirely in fp32 (order='f32')."""
    out = []
    for m, s, e in zip(mean, std, eps):
        if order == "f64":
            w = np.asarray(m, np.float64) + (np.float64(inflate) * np.asarray(s, np.float64)) * np.asarray(e, np.float64)
        else:
            w = np.asarray(m, np.float32) + (np.float32(inflate) * np.asarray(s, np.float32)) * np.asarray(e, np.float32)
        out.append(w.astype(np.float32))
    return out


def sample_reference_rng(mean, std, inflate=1.0):
    """Exactly what the reference executes (global NumPy RNG, tensor order W0,b0,W1,b1,...):
    posteriormodel.py:70-75."""
    return [np.random.normal(loc=mean[i], scale=inflate * std[i]) for i in range(len(mean))]


def draw_stored_index(freq):
    """posteriormodel.py:77: np.random.choice(range(S), p=frequency)."""
    return int(np.random.choice(range(len(freq)), p=freq))


# --------------------------------------------------------------------------------------
# predict / predict_logits (a3, a4)
# --------------------------------------------------------------------------------------


def predict(layers, mean, std, x, n, eps=None, inflate=1.0, out="probs", bf16=False, order="f64"):
    """PosteriorModel.predict (posteriormodel.py:81-101) / predict_logits (:114-139):
    fp32 running sum over samples in sample order, then ``/ float(n)``.
    ``eps``: list (len n) of per-tensor arrays; None -> reference RNG (np.random.normal)."""
    acc = None
    for s in range(n):
        if eps is None:
            w = [np.asarray(t, np.float32) for t in sample_reference_rng(mean, std, inflate)]
        else:
            w = sample_gaussian(mean, std, eps[s], inflate, order)
        y = forward(layers, w, x, np.float32, out, bf16)
        acc = y if acc is None else acc + y  # posteriormodel.py:97-100
    return (acc / np.float32(float(n))).astype(np.float32)  # :101


def predict_moments(layers, mean, std, x, n, eps, inflate=1.0, bf16=False, order="f64"):
    """Sum p and sum p^2 over samples (uncertainty.py:15-36 consumes per-sample predictions;
    the GPU epilogue emits the two running sums instead of the [n,B,C] stack)."""
    s1 = s2 = None
    for s in range(n):
        w = sample_gaussian(mean, std, eps[s], inflate, order)
        y = forward(layers, w, x, np.float32, "probs", bf16)
        s1 = y if s1 is None else s1 + y
        s2 = y * y if s2 is None else s2 + y * y
    return s1, s2


def variational_uncertainty(per_sample_preds):
    """uncertainty.py:33-35."""
    p_hat = np.asarray(per_sample_preds)
    epistemic = np.mean(p_hat ** 2, axis=0) - np.mean(p_hat, axis=0) ** 2
    aleatoric = np.mean(p_hat * (1 - p_hat), axis=0)
    return epistemic, aleatoric


def predict_stored(layers, samples, x, idx, weights=None, out="probs", bf16=False):
    """Sample-based posterior (HMC/SGHMC store, hmc.py:308-322).  ``idx`` = the stored-sample
    indices visited in order (posteriormodel.py:77-78 draws them with replacement ~ freq);
    ``weights`` = per-visit weights (1 for the MC loop; freq_i for the deterministic
    expectation sum_i freq_i f(W_i) used by probverification.py:619-650).  Returns the SUM."""
    acc = None
    for j, i in enumerate(idx):
        y = forward(layers, samples[i], x, np.float32, out, bf16)
        if weights is not None:
            y = y * np.float32(weights[j])
        acc = y if acc is None else acc + y
    return acc


def normalise_freq(freq):
    """posteriormodel.py:52-54."""
    freq = np.asarray(freq, dtype=np.float64)
    return freq / np.sum(freq)


# --------------------------------------------------------------------------------------
# Bayes-by-Backprop (a7, a8, a9, a10)
# --------------------------------------------------------------------------------------


def bbb_forward(layers, mean, rho, eps, x, bf16=False):
    """bayesbybackprop.py:50-65: w = mean + softplus(rho) * noise (fp32), forward.
    Returns (predictions, sampled_weights)."""
    w = []
    for m, r, e in zip(mean, rho, eps):
        w.append((np.asarray(m, np.float32) + softplus(r) * np.asarray(e, np.float32)).astype(np.float32))
    return forward(layers, w, x, np.float32, "probs", bf16), w


def implicit_prior(layers, inflate_prior=1.0):
    """optimizer.py:204-228: zero mean, std = sqrt(inflate_prior / fan_in), fan_in =
    prod(shape[:-1]) for conv kernels, bias gets its kernel's std."""
    mean, var = [], []
    for l in layers:
        shp = l.weight_shapes()
        if not shp:
            continue
        sha, b_sha = shp
        nin = int(np.prod(sha[:-1])) if len(sha) > 2 else sha[0]
        std = math.sqrt(inflate_prior / nin)
        mean += [np.zeros(sha, np.float32), np.zeros(b_sha, np.float32)]
        var += [np.ones(sha, np.float32) * np.float32(std), np.ones(b_sha, np.float32) * np.float32(std)]
    return mean, var


def prior_generator(layers, means, vars_, inflate_prior=1.0):
    """optimizer.py:230-257."""
    if type(means) == int and type(vars_) == int and (means < 0 or vars_ < 0):
        return implicit_prior(layers, inflate_prior)
    shapes = weight_shapes(layers)
    if type(means) in (int, float):
        means = [0.0 if means == -1 else means] * len(shapes)
    if type(vars_) in (int, float):
        vars_ = [0.0 if vars_ == -1 else vars_] * len(shapes)
    m, v = [], []
    for i, s in enumerate(shapes):
        pi = i // 2  # optimizer.py:251 floor(index/2)
        m.append(np.ones(s, np.float32) * np.float32(means[pi]))
        v.append(np.ones(s, np.float32) * np.float32(vars_[pi]))
    return m, v


# --------------------------------------------------------------------------------------
# Statistical verification outer loop (a15)
# --------------------------------------------------------------------------------------


def okamoto_bound(epsilon, delta):
    """verifiers.py:517-518."""
    return (-1 * .5) * math.log(float(delta) / 2) * (1.0 / (epsilon ** 2))


def chernoff_n(confidence=0.95, delta=0.3):
    """verifiers.py:542-543 / :579-580."""
    epsilon = 1 - confidence
    return math.ceil((1 / (2 * epsilon ** 2)) * math.log(2 / delta))


def absolute_massart_halting(succ, trials, I, epsilon, delta, alpha):
    """verifiers.py:521-532."""
    if I[0] < 0.5 and I[1] > 0.5:
        return -1
    elif I[1] < 0.5:
        val = I[1]
        h = (9 / 2.0) * (((3 * val + epsilon) * (3 * (1 - val) - epsilon)) ** (-1))
        return math.ceil((h * (epsilon ** 2)) ** (-1) * math.log((delta - alpha) ** (-1)))
    elif I[0] >= 0.5:
        val = I[0]
        h = (9 / 2.0) * (((3 * (1 - val) + epsilon) * ((3 * val) + epsilon)) ** (-1))
        return math.ceil((h * (epsilon ** 2)) ** (-1) * math.log((delta - alpha) ** (-1)))


def clopper_pearson(successes, trials, alpha=0.05):
    """statsmodels.stats.proportion.proportion_confint(method='beta') as called at
    verifiers.py:636 (third-party, restated from its published definition), with the NaN
    handling of verifiers.py:637-640."""
    from scipy.stats import beta

    lb = beta.ppf(alpha / 2, successes, trials - successes + 1)
    ub = beta.ppf(1 - alpha / 2, successes + 1, trials - successes)
    lb = 0.0 if math.isnan(lb) else float(lb)
    ub = 1.0 if math.isnan(ub) else float(ub)
    return lb, ub


def count_predicate(layers, mean, std, x, labels, n, eps, inflate=1.0, bf16=False, order="f64"):
    """BASELINE config 5 restated: per posterior sample, forward the K fixed (pre-perturbed)
    inputs and apply the tutorial predicate ``argmax == TRUE_VALUE`` (verifiers.py:628-634
    with the notebook's predicate); returns flags [n,K] uint8 and counts [K]."""
    K = np.asarray(x).shape[0]
    flags = np.zeros((n, K), np.uint8)
    for s in range(n):
        w = sample_gaussian(mean, std, eps[s], inflate, order)
        z = forward(layers, w, x, np.float32, "logits", bf16).reshape(K, -1)
        flags[s] = (np.argmax(z, axis=-1) == np.asarray(labels).reshape(K)).astype(np.uint8)
    return flags, flags.sum(axis=0).astype(np.int64)


def massart_loop(flags_1d, confidence=0.95, delta=0.3, alpha=0.05):
    """The sequential halting logic of massart_bound_check (verifiers.py:578-653) applied to
    a pre-computed 0/1 success sequence (one entry per posterior sample, in order).
    Returns (successes/iterations, iterations)."""
    epsilon = 1 - confidence
    chernoff_bound = chernoff_n(confidence, delta)
    successes, iterations = 0.0, 0.0
    halting_bound = chernoff_bound
    while iterations <= halting_bound:
        successes += float(flags_1d[int(iterations)])
        iterations += 1
        lb, ub = clopper_pearson(successes, iterations)
        hb = absolute_massart_halting(successes, iterations, [lb, ub], epsilon, delta, alpha)
        halting_bound = chernoff_bound if hb == -1 else min(hb, chernoff_bound)
    return successes / iterations, iterations


# --------------------------------------------------------------------------------------
# Philox4x32-10 counter RNG (the product's in-kernel generator, restated for bit checks)
# --------------------------------------------------------------------------------------

_PH_M0, _PH_M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_PH_W0, _PH_W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)


# ---------------------------------------------------------------------------------------
# Interval bound propagation over sampled weights
# ---------------------------------------------------------------------------------------
def propagate_interval(W, b, x_l, x_u, marg=0.0, b_marg=0.0):
    """One Dense layer on an interval, centre/radius form in float64 (verifiers.py:23-41):
    h_mu = x_mu W, x_rad = x_r |W|, W_rad = |x_mu| W_r, Quad = |x_r| |W_r|;
    h_u = h_mu + x_rad + W_rad + Quad + (b + b_marg), h_l = h_mu - x_rad - W_rad - Quad + (b - b_marg).
    (x_u + x_l)/2 and (x_u - x_l)/2 are formed in the dtype of the incoming bounds and THEN cast (:26-27)."""
    x_l, x_u = np.asarray(x_l), np.asarray(x_u)
    x_mu = ((x_u + x_l) / 2).astype(np.float64)
    x_r = ((x_u - x_l) / 2).astype(np.float64)
    W_mu = np.asarray(W).astype(np.float64)
    W_r = np.zeros_like(W_mu) if (np.isscalar(marg) and marg == 0) else np.broadcast_to(np.asarray(marg, np.float64), W_mu.shape)
    b_u = (np.asarray(b) + b_marg).astype(np.float64)
    b_l = (np.asarray(b) - b_marg).astype(np.float64)
    h_mu = x_mu @ W_mu
    x_rad = x_r @ np.abs(W_mu)
    W_rad = np.abs(x_mu) @ W_r
    quad = np.abs(x_r) @ np.abs(W_r)
    h_u = h_mu + x_rad + W_rad + quad + b_u
    h_l = h_mu - x_rad - W_rad - quad + b_l
    return h_l, h_u


def _conv_valid(x, k):
    """tf.nn.convolution(x, k) with its defaults: NHWC x HWIO, VALID padding, stride 1."""
    kh, kw, cin, cout = k.shape
    n, h, w, c = x.shape
    ho, wo = h - kh + 1, w - kw + 1
    cols = np.empty((n, ho, wo, kh, kw, c), dtype=np.result_type(x.dtype, k.dtype))
    for i in range(kh):
        for j in range(kw):
            cols[:, :, :, i, j, :] = x[:, i:i + ho, j:j + wo, :]
    return (cols.reshape(n * ho * wo, kh * kw * c) @ k.reshape(kh * kw * c, cout)).reshape(n, ho, wo, cout)


def propagate_conv2d(W, b, x_l, x_u, marg=0, b_marg=0):
    """One Conv2D layer on an interval, exactly as verifiers.py:11-21 writes it:
      w_pos = max(W + marg, 0), w_neg = min(W - marg, 0)
      h_l = conv(x_l, w_pos) + conv(x_u, w_neg);  h_u = conv(x_u, w_pos) + conv(x_l, w_neg)
      nom = conv((x_l + x_u) / 2, W);  h_l = nom + h_l + (b - b_marg);  h_u = nom + h_u + (b + b_marg)
    i.e. the nominal response is ADDED to both interval sums (kept: it is what the reference computes), and
    tf.nn.convolution runs with its defaults (VALID, stride 1) whatever the layer's padding / strides are.
    Arithmetic in float64 here; TensorFlow computes it in the inputs' float32."""
    W = np.asarray(W, np.float64)
    x_l, x_u = np.asarray(x_l, np.float64), np.asarray(x_u, np.float64)
    w_pos, w_neg = np.maximum(W + marg, 0), np.minimum(W - marg, 0)
    h_l = _conv_valid(x_l, w_pos) + _conv_valid(x_u, w_neg)
    h_u = _conv_valid(x_u, w_pos) + _conv_valid(x_l, w_neg)
    nom = _conv_valid((x_l + x_u) / 2, W)
    return nom + h_l + (np.asarray(b, np.float64) - b_marg), nom + h_u + (np.asarray(b, np.float64) + b_marg)


def ibp_forward(layers, weights, x, eps, predict=False):
    """`IBP(model, inp, weights, eps, predict)` (verifiers.py:68-95) for Dense / Conv2D / weightless layers:
    input box clip(x -+ eps, 0, 1) (predict=False) or x -+ eps; every Dense through
    propagate_interval; the layer's activation on both bounds, except that with predict=False
    the LAST layer's activation is skipped (:90-91), i.e. the result is a logit interval.
    Returns (h_l, h_u) float64."""
    x = np.asarray(x)
    if not predict:
        h_u = np.clip(x + eps, 0.0, 1.0)
        h_l = np.clip(x - eps, 0.0, 1.0)
    else:
        h_u, h_l = x + eps, x - eps
    wi = 0
    for i, l in enumerate(layers):
        if l.kind == "flatten":
            h_u, h_l = h_u.reshape(h_u.shape[0], -1), h_l.reshape(h_l.shape[0], -1)
            continue
        if l.kind == "maxpool2d":          # a weightless layer is applied to each bound (:78-81)
            h_u, h_l = _maxpool(h_u, l), _maxpool(h_l, l)
            continue
        if l.kind == "conv2d":
            h_l, h_u = propagate_conv2d(weights[wi], weights[wi + 1], h_l, h_u)
        else:
            h_l, h_u = propagate_interval(weights[wi], weights[wi + 1], h_l, h_u)
        wi += 2
        if (not predict) and i >= len(layers) - 1:
            continue
        h_l, h_u = _act(l.activation, h_l), _act(l.activation, h_u)
    return h_l, h_u


def ibp_state(layers, weights, s0, s1, sigma, weight_margin=0.0):
    """`IBP_state(model, s0, s1, weights, weight_margin, logits=True)` (verifiers.py:97-123 = probverification.IBP_prob
    :84-110): the input box [s0, s1] as given, marg = weight_margin * posterior_var per tensor (:107-110), every Dense
    through propagate_interval with the |x_mu| W_r and |x_r| |W_r| terms, activation on all but the last layer."""
    h_l, h_u = np.asarray(s0), np.asarray(s1)
    wi = 0
    for i, l in enumerate(layers):
        if l.kind == "flatten":
            h_u, h_l = h_u.reshape(h_u.shape[0], -1), h_l.reshape(h_l.shape[0], -1)
            continue
        if l.kind != "dense":
            raise NotImplementedError("oracle IBP covers Dense / Flatten")
        marg = weight_margin * np.asarray(sigma[wi])          # in sigma's dtype (fp32), like the reference (:109-110)
        b_marg = weight_margin * np.asarray(sigma[wi + 1])
        h_l, h_u = propagate_interval(weights[wi], weights[wi + 1], h_l, h_u, marg=marg, b_marg=b_marg)
        wi += 2
        if i < len(layers) - 1:
            h_l, h_u = _act(l.activation, h_l), _act(l.activation, h_u)
    return h_l, h_u


def worst_case_logits(logit_l, logit_u, cls, depth=None):
    """v1 * logit_l + v2 * logit_u with v1 = one_hot(cls), v2 = 1 - v1 (verifiers.py:549-552, :590-593):
    the true class takes its lower bound, every other class its upper bound."""
    logit_l, logit_u = np.asarray(logit_l), np.asarray(logit_u)
    C = logit_l.shape[-1] if depth is None else depth
    v1 = np.eye(C, dtype=logit_l.dtype)[np.asarray(cls).reshape(-1)]
    v1 = v1.reshape(logit_l.shape[:-1] + (C,)) if v1.shape[0] != 1 else np.broadcast_to(v1[0], logit_l.shape)
    return (1 - v1) * logit_u + v1 * logit_l


def chernoff_verification(layers, mean, std, x, eps, cls, n, noise, inflate=1.0, order="f64"):
    """`chernoff_bound_verification` (verifiers.py:537-557) with the weight noise injected: for each of
    the n posterior samples, IBP logits on the eps-ball, worst case w.r.t. cls, the last layer's
    activation, and the SUM over samples (the reference returns the running sum, not the mean).
    Also returns the per-sample predicate argmax(worst case) == cls that massart_bound_check
    (:588-634, verify=True, classification=True) counts."""
    last_act = layers[-1].activation
    total = None
    flags = []
    for s in range(n):
        w = sample_gaussian(mean, std, noise[s], inflate, order)
        lo, hi = ibp_forward(layers, w, x, eps, predict=False)
        wc = worst_case_logits(lo, hi, cls)
        sm = _act(last_act, wc)
        total = sm if total is None else total + sm
        flags.append(np.argmax(wc, axis=-1) == np.asarray(cls).reshape(-1))
    return total, np.asarray(flags)



# ---------------------------------------------------------------------------------------
# Bayes-by-Backprop training step
# ---------------------------------------------------------------------------------------
_LOG_SQRT_2PI = 0.5 * np.log(2.0 * np.pi)


def log_q(x, mu, rho, dtype=np.float64):
    """losses.py:10-16: sum over tensors of the MEAN over elements of Normal(mu_i, softplus(rho_i)).log_prob(x_i)."""
    total = dtype(0)
    for xi, mi, ri in zip(x, mu, rho):
        sg = softplus(np.asarray(ri, dtype), dtype)
        z = (np.asarray(xi, dtype) - np.asarray(mi, dtype)) / sg
        total = total + np.mean(-0.5 * z * z - np.log(sg) - dtype(_LOG_SQRT_2PI))
    return total


def sparse_categorical_crossentropy(labels, probs, dtype=np.float64):
    """tf.keras.losses.SparseCategoricalCrossentropy() on probabilities: mean over the batch of -log(clip(p[label], 1e-7, 1-1e-7))."""
    p = np.asarray(probs, dtype)
    pl = p[np.arange(p.shape[0]), np.asarray(labels).reshape(-1)]
    return -np.mean(np.log(np.clip(pl, dtype(1e-7), dtype(1 - 1e-7))))


def kl_loss(layers, labels, x, W, prior_mean, prior_var, q_mean, q_rho, kl_weight=1.0, dtype=np.float64):
    """losses.KL_Loss (losses.py:40-45): data_likli + kl_weight * (log_q(W; q_mean, q_rho) - log_q(W; prior_mean, prior_var)),
    and the KL component.  NOTE the prior goes through the same log_q, i.e. its scale is softplus(prior_var)."""
    pred = forward(layers, [np.asarray(w, dtype) for w in W], np.asarray(x, dtype), dtype=dtype)
    data = sparse_categorical_crossentropy(labels, pred, dtype)
    klc = log_q(W, q_mean, q_rho, dtype) - log_q(W, prior_mean, prior_var, dtype)
    return data + dtype(kl_weight) * klc, klc, pred


def _mlp_forward_acts(dense, W, x):
    acts, h = [x], x
    for li, l in enumerate(dense):
        h = _act(l.activation, h @ W[2 * li] + W[2 * li + 1])
        acts.append(h)
    return acts


def _mlp_backward_weights(dense, W, acts, dz):
    """dLoss/dW for every tensor given dLoss/d(logits) = dz; acts[i] = input of layer i, acts[-1] = output."""
    gW = [None] * len(W)
    for li in range(len(dense) - 1, -1, -1):
        gW[2 * li] = acts[li].T @ dz
        gW[2 * li + 1] = dz.sum(0)
        if li > 0:
            da = dz @ W[2 * li].T
            a, name = acts[li], dense[li - 1].activation
            if name == "relu":
                dz = da * (a > 0)
            elif name == "tanh":
                dz = da * (1 - a * a)
            elif name == "sigmoid":
                dz = da * a * (1 - a)
            elif name in ("linear", None):
                dz = da
            else:
                raise NotImplementedError(name)
    return gW


def _mlp_forward_acts_bf16(dense, W, x):
    """_mlp_forward_acts as the GPU's bf16 training step computes it: GEMM operands (input, kernels, stored hidden outputs)
    rounded to bf16, fp32 accumulation / bias / activation; the last entry (the softmax output) stays fp32."""
    h = bf16_round(x)
    acts = [h]
    for li, l in enumerate(dense):
        h = _act(l.activation, (h @ bf16_round(W[2 * li]) + W[2 * li + 1]).astype(np.float32))
        if li + 1 < len(dense):
            h = bf16_round(h)
        acts.append(h)
    return acts


def _mlp_backward_weights_bf16(dense, W, acts, dz):
    """_mlp_backward_weights with the same roundings: the upstream gradient of every layer is a bf16 GEMM operand."""
    gW = [None] * len(W)
    dz = bf16_round(dz)
    for li in range(len(dense) - 1, -1, -1):
        gW[2 * li] = (acts[li].T.astype(np.float64) @ dz.astype(np.float64)).astype(np.float32)
        gW[2 * li + 1] = dz.astype(np.float64).sum(0).astype(np.float32)
        if li > 0:
            da = (dz.astype(np.float64) @ bf16_round(W[2 * li]).T.astype(np.float64)).astype(np.float32)
            dz = bf16_round(da * _act_grad(dense[li - 1].activation, acts[li]))
    return gW


def _act_grad(name, a):
    """Derivative of an element-wise activation written in terms of its OUTPUT a."""
    if name == "relu":
        return (a > 0).astype(a.dtype)
    if name == "tanh":
        return 1 - a * a
    if name == "sigmoid":
        return a * (1 - a)
    if name in ("linear", None):
        return np.ones_like(a)
    raise NotImplementedError(name)


def _ibp_forward_bounds(dense, W, x, eps):
    """verifiers.IBP (:68-95, predict=False) on a Dense MLP keeping every layer's bounds: U[i] / Lo[i] = upper / lower
    input of layer i (after the previous activation), U[-1] / Lo[-1] = the logit interval.  Also returns the pre-activation
    bounds.  The input box is clip(x -+ eps, 0, 1) (:70-71)."""
    hu, hl = np.clip(x + eps, 0.0, 1.0), np.clip(x - eps, 0.0, 1.0)
    U, Lo, Z = [hu], [hl], []
    for li, l in enumerate(dense):
        zl, zu = propagate_interval(W[2 * li], W[2 * li + 1], hl, hu)
        Z.append((zl, zu))
        if li < len(dense) - 1:
            hl, hu = _act(l.activation, zl), _act(l.activation, zu)
        else:
            hl, hu = zl, zu
        U.append(hu); Lo.append(hl)
    return U, Lo, Z


def _ibp_backward_weights(dense, W, U, Lo, dzu, dzl):
    """dLoss/dW through the interval forward, given dLoss/d(upper logits) = dzu and dLoss/d(lower logits) = dzl.
    With mu = (u + l)/2, r = (u - l)/2 of a layer's input and MU = mu W + b, R = r |W| (propagate_interval,
    verifiers.py:23-41): z_u = MU + R, z_l = MU - R, so dMU = dzu + dzl, dR = dzu - dzl,
    dW = mu^T dMU + sign(W) * (r^T dR), db = sum_b dMU, dmu = dMU W^T, dr = dR |W|^T, du = (dmu + dr)/2, dl = (dmu - dr)/2."""
    gW = [None] * len(W)
    for li in range(len(dense) - 1, -1, -1):
        mu, r = (U[li] + Lo[li]) / 2, (U[li] - Lo[li]) / 2
        dMU, dR = dzu + dzl, dzu - dzl
        gW[2 * li] = mu.T @ dMU + np.sign(W[2 * li]) * (r.T @ dR)
        gW[2 * li + 1] = dMU.sum(0)
        if li > 0:
            dmu, dr = dMU @ W[2 * li].T, dR @ np.abs(W[2 * li]).T
            name = dense[li - 1].activation
            dzu = (dmu + dr) / 2 * _act_grad(name, U[li])
            dzl = (dmu - dr) / 2 * _act_grad(name, Lo[li])
    return gW


def bbb_step(layers, mean, rho, prior_mean, prior_var, x, labels, noise, lrate, kl_weight=1.0, dtype=np.float64, bf16=False,
             robust_train=0, epsilon=0.0, robust_lambda=1.0, lower=-np.inf, upper=np.inf, eps_draws=None):
    """`BayesByBackprop.step` (bayesbybackprop.py:46-170) for Dense MLPs with a softmax output and the sparse categorical
    cross-entropy, gradients written out by hand (the reference gets them from tf.GradientTape):
      W = mean + softplus(rho) * noise                                                    (:50-59)
      robust_train == 0: output = model_W(x)                                               (:65-72)
      robust_train == 1: (l, u) = IBP(x, W, eps), worst = softmax(u off the label, l on it),
                         output = lambda * model_W(x) + (1 - lambda) * worst                (:73-90)
      robust_train == 3: output = mean over eps_draws of softmax(worst-case IBP logits at that eps)      (:102-121)
      robust_train == 2: x_adv = FGSM(x, eps) against W towards argmax model_W(x) (:92, attacks.py:81-102 with the
                         model's current weights), output = lambda * model_W(x) + (1 - lambda) * model_W(x_adv)   (:93-95)
      weight_gradient = dLoss/dW       (data term + kl_weight * d(log_q - log_prior)/dW)  (:138)
      mean_gradient   = dLoss/dmean = kl_weight * (W - mean) / (sigma^2 n_i)              (:139)
      var_gradient    = dLoss/drho  = kl_weight * ((W - mean)^2 / sigma^3 - 1 / sigma) * sigmoid(rho) / n_i   (:140)
      posti_mean_grad = weight_gradient + mean_gradient                                   (:146)
      posti_var_grad  = noise / (1 + exp(-rho)) * weight_gradient + var_gradient          (:148-150)
      new = old - lrate * grad                                                            (:156-163)
    Returns (new_mean, new_rho, loss, kl_component, predictions, W)."""
    d = dtype
    mean = [np.asarray(t, d) for t in mean]
    rho = [np.asarray(t, d) for t in rho]
    noise = [np.asarray(t, d) for t in noise]
    pm = [np.asarray(t, d) for t in prior_mean]
    ps = [softplus(np.asarray(t, d), d) for t in prior_var]
    sg = [softplus(r, d) for r in rho]
    W = [m + s * e for m, s, e in zip(mean, sg, noise)]
    dense = [l for l in layers if l.kind == "dense"]
    if len(dense) != len(layers):
        raise NotImplementedError("oracle bbb_step covers Dense MLPs")
    if dense[-1].activation != "softmax":
        raise NotImplementedError("oracle bbb_step: softmax output")
    x = np.asarray(x, d)
    if bf16:
        # the GPU's bf16 training mode: the GEMMs see bf16 operands (dtype should be np.float32 so that W is the fp32 draw the
        # device rounds); draw, losses, KL terms and the update are unchanged
        if robust_train not in (0, 2):
            raise NotImplementedError("bf16 emulation: robust_train 0 and 2")
        fwd, bwd = _mlp_forward_acts_bf16, _mlp_backward_weights_bf16
    else:
        fwd, bwd = _mlp_forward_acts, _mlp_backward_weights
    acts = fwd(dense, W, x)
    pred = acts[-1]
    B = pred.shape[0]
    lab = np.asarray(labels).reshape(-1)
    onehot = np.zeros_like(pred)
    onehot[np.arange(B), lab] = 1
    if robust_train == 0:
        out = pred
        pl = out[np.arange(B), lab]
        active = ((pl > 1e-7) & (pl < 1 - 1e-7)).astype(d)
        gW = bwd(dense, W, acts, (pred - onehot) * active[:, None] / d(B))
    elif robust_train == 2:
        direc = pred.argmax(1)
        x_adv = fgsm(x, input_gradient(layers, [W], x, direc, d, bf16=bf16), epsilon, lower, upper)
        acts_adv = fwd(dense, W, x_adv)
        pa = acts_adv[-1]
        out = d(robust_lambda) * pred + d(1 - robust_lambda) * pa
        oy = out[np.arange(B), lab]
        g = np.where((oy > 1e-7) & (oy < 1 - 1e-7), -1.0 / (B * oy), 0.0)
        dz = (d(robust_lambda) * g * pred[np.arange(B), lab])[:, None] * (onehot - pred)
        dza = (d(1 - robust_lambda) * g * pa[np.arange(B), lab])[:, None] * (onehot - pa)
        g1 = bwd(dense, W, acts, dz)
        g2 = bwd(dense, W, acts_adv, dza)
        gW = [a + b for a, b in zip(g1, g2)]
    elif robust_train == 1:
        U, Lo, _ = _ibp_forward_bounds(dense, W, x, epsilon)
        worst = (1 - onehot) * U[-1] + onehot * Lo[-1]                     # :78-82 (v2 * logit_u + v1 * logit_l)
        pw = _act("softmax", worst)                                        # :84
        out = d(robust_lambda) * pred + d(1 - robust_lambda) * pw          # :85
        oy = out[np.arange(B), lab]
        g = np.where((oy > 1e-7) & (oy < 1 - 1e-7), -1.0 / (B * oy), 0.0)
        dz = (d(robust_lambda) * g * pred[np.arange(B), lab])[:, None] * (onehot - pred)
        dzw = (d(1 - robust_lambda) * g * pw[np.arange(B), lab])[:, None] * (onehot - pw)
        g1 = _mlp_backward_weights(dense, W, acts, dz)
        g2 = _ibp_backward_weights(dense, W, U, Lo, dzw * (1 - onehot), dzw * onehot)
        gW = [a + b for a, b in zip(g1, g2)]
    elif robust_train == 3:
        # :102-121 -- output = sum_k (1/K) softmax(worst-case IBP logits at eps_k); eps_k ~ Exponential(1 / epsilon) in the
        # reference, injected here (eps_draws); the clean prediction is not part of the loss
        K = len(eps_draws)
        bounds, pws = [], []
        out = np.zeros_like(pred)
        for ek in eps_draws:
            U, Lo, _ = _ibp_forward_bounds(dense, W, x, d(ek))
            pw = _act("softmax", (1 - onehot) * U[-1] + onehot * Lo[-1])
            bounds.append((U, Lo)); pws.append(pw)
            out = out + d(1.0 / K) * pw
        oy = out[np.arange(B), lab]
        g = np.where((oy > 1e-7) & (oy < 1 - 1e-7), -1.0 / (B * oy), 0.0)
        gW = None
        for (U, Lo), pw in zip(bounds, pws):
            dzw = (d(1.0 / K) * g * pw[np.arange(B), lab])[:, None] * (onehot - pw)
            gk = _ibp_backward_weights(dense, W, U, Lo, dzw * (1 - onehot), dzw * onehot)
            gW = gk if gW is None else [a + b for a, b in zip(gW, gk)]
    else:
        raise NotImplementedError("oracle bbb_step: robust_train 0, 1, 2 and 3")
    data = sparse_categorical_crossentropy(lab, out, d)
    klc = log_q(W, mean, rho, d) - log_q(W, pm, [np.asarray(t, d) for t in prior_var], d)
    loss = data + d(kl_weight) * klc
    new_mean, new_rho = [], []
    for i in range(len(W)):
        n_i = d(W[i].size)
        sig = 1.0 / (1.0 + np.exp(-rho[i]))
        wg = gW[i] + d(kl_weight) * (-(W[i] - mean[i]) / (sg[i] ** 2) + (W[i] - pm[i]) / (ps[i] ** 2)) / n_i
        mg = d(kl_weight) * (W[i] - mean[i]) / (sg[i] ** 2) / n_i
        vg = d(kl_weight) * (((W[i] - mean[i]) ** 2) / sg[i] ** 3 - 1.0 / sg[i]) * sig / n_i
        g_mean = wg + mg
        g_rho = noise[i] * sig * wg + vg
        new_mean.append(mean[i] - d(lrate) * g_mean)
        new_rho.append(rho[i] - d(lrate) * g_rho)
    return new_mean, new_rho, loss, klc, pred, W


# ---------------------------------------------------------------------------------------
# Expected input gradient and the attacks built on it
# ---------------------------------------------------------------------------------------
def mean_prediction_loss(layers, weight_sets, x, labels, dtype=np.float64):
    """loss_fn(direction, model.predict(inp)) of attacks.py:64-65 for a PosteriorModel: the sparse categorical
    cross-entropy of the MEAN over the given weight sets of the model output (posteriormodel.py:94-101)."""
    pbar = None
    for w in weight_sets:
        p = forward(layers, [np.asarray(t, dtype) for t in w], np.asarray(x, dtype), dtype=dtype)
        pbar = p if pbar is None else pbar + p
    pbar = pbar / dtype(len(weight_sets))
    return sparse_categorical_crossentropy(labels, pbar, dtype), pbar


def input_gradient(layers, weight_sets, x, labels, dtype=np.float64, bf16=False):
    """d mean_prediction_loss / d x by hand (the reference uses tf.GradientTape, attacks.py:60-67) for Dense MLPs with
    a softmax output: with pbar the mean prediction, dz_s[b, c] = g_b p_s[b, y_b] ([c == y_b] - p_s[b, c]),
    g_b = -1 / (B n pbar_b[y_b]), back-propagated through each sample's layers and summed over samples.
    ``bf16=True`` emulates the GPU's bf16 mode: every GEMM operand (x, weights, stored hidden activations, the upstream
    gradients between layers) rounded to bf16, fp32 accumulation, bias / softmax / cross-entropy / sums in fp32."""
    d = np.float32 if bf16 else dtype
    rnd = bf16_round if bf16 else (lambda a: a)
    x = rnd(np.asarray(x, d))
    lab = np.asarray(labels).reshape(-1)
    dense = [l for l in layers if l.kind == "dense"]
    B, n = x.shape[0], len(weight_sets)
    fw = []
    for w in weight_sets:
        W = [np.asarray(t, d) for t in w]
        if bf16:
            W = [rnd(t) if i % 2 == 0 else t for i, t in enumerate(W)]
        acts = [x]
        h = x
        for li, l in enumerate(dense):
            h = _act(l.activation, (h @ W[2 * li] + W[2 * li + 1]).astype(d))
            if li + 1 < len(dense):
                h = rnd(h)
            acts.append(h)
        fw.append((W, acts))
    pbar = sum(a[-1] for _, a in fw) / d(n)
    py = pbar[np.arange(B), lab]
    g = np.where((py > 1e-7) & (py < 1 - 1e-7), -1.0 / (B * n * py), 0.0).astype(d)
    grad = np.zeros_like(x)
    for W, acts in fw:
        p = acts[-1]
        onehot = np.zeros_like(p)
        onehot[np.arange(B), lab] = 1
        dz = rnd((g * p[np.arange(B), lab])[:, None] * (onehot - p))
        for li in range(len(dense) - 1, -1, -1):
            da = dz @ W[2 * li].T
            if li == 0:
                grad += da
                break
            a, name = acts[li], dense[li - 1].activation
            dz = rnd(da * ((a > 0) if name == "relu" else (1 - a * a) if name == "tanh" else a * (1 - a) if name == "sigmoid" else 1.0))
    return grad


def fgsm(x, grad, eps, lower=-np.inf, upper=np.inf):
    """attacks.py:94-101: adv = clip(clip(x + eps * sign(grad), x - eps, x + eps), input_lower, input_upper)."""
    x = np.asarray(x)
    adv = np.clip(x + eps * np.sign(grad), x - eps, x + eps)
    return np.clip(adv, lower, upper)



def philox4x32_10(ctr, key):
    """Philox4x32-10 (Salmon et al., SC'11) on arrays: ctr [..,4] uint32, key [..,2] uint32."""
    c = [np.asarray(ctr[..., i], np.uint32).copy() for i in range(4)]
    k0 = np.asarray(key[..., 0], np.uint32).copy()
    k1 = np.asarray(key[..., 1], np.uint32).copy()
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = _PH_M0 * c[0].astype(np.uint64)
            p1 = _PH_M1 * c[2].astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), p0.astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), p1.astype(np.uint32)
            c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
            k0 = (k0 + _PH_W0).astype(np.uint32)
            k1 = (k1 + _PH_W1).astype(np.uint32)
    return np.stack(c, axis=-1)


def philox_normals(seed: int, sample: int, P: int) -> np.ndarray:
    """The eps stream the CUDA library generates for posterior sample ``sample``:
    element j (flat Keras order over all tensors) comes from Philox block j//4, lane j%4;
    counter = (j//4 lo, j//4 hi, sample lo, sample hi), key = (seed lo, seed hi);
    lanes (0,1) and (2,3) feed one Box-Muller each:
      u1 = (r_a + 1) * 2^-32 in (0,1], u2 = r_b * 2^-32 in [0,1),
      z_a = sqrt(-2 ln u1) cos(2 pi (u2-1/2)), z_b = sqrt(-2 ln u1) sin(2 pi (u2-1/2))
    (angle kept in [-pi, pi) where the GPU's fast sincos is most accurate)."""
    nb = (P + 3) // 4
    blk = np.arange(nb, dtype=np.uint64)
    ctr = np.stack([
        (blk & np.uint64(0xFFFFFFFF)).astype(np.uint32), (blk >> np.uint64(32)).astype(np.uint32),
        np.full(nb, sample & 0xFFFFFFFF, np.uint32), np.full(nb, (sample >> 32) & 0xFFFFFFFF, np.uint32)], axis=-1)
    key = np.stack([np.full(nb, seed & 0xFFFFFFFF, np.uint32), np.full(nb, (seed >> 32) & 0xFFFFFFFF, np.uint32)], axis=-1)
    r = philox4x32_10(ctr, key).astype(np.float64)
    z = np.empty((nb, 4), np.float64)
    for a, b in ((0, 1), (2, 3)):
        u1 = (r[:, a] + 1.0) * 2.0 ** -32
        u2 = r[:, b] * 2.0 ** -32
        rad = np.sqrt(-2.0 * np.log(u1))
        z[:, a] = rad * np.cos(2 * np.pi * (u2 - 0.5))
        z[:, b] = rad * np.sin(2 * np.pi * (u2 - 0.5))
    return z.reshape(-1)[:P].astype(np.float32)


def philox_raw(seed: int, sample: int, P: int) -> np.ndarray:
    """Raw uint32 outputs matching philox_normals' blocks (bit-exact check of the kernel)."""
    nb = (P + 3) // 4
    blk = np.arange(nb, dtype=np.uint64)
    ctr = np.stack([
        (blk & np.uint64(0xFFFFFFFF)).astype(np.uint32), (blk >> np.uint64(32)).astype(np.uint32),
        np.full(nb, sample & 0xFFFFFFFF, np.uint32), np.full(nb, (sample >> 32) & 0xFFFFFFFF, np.uint32)], axis=-1)
    key = np.stack([np.full(nb, seed & 0xFFFFFFFF, np.uint32), np.full(nb, (seed >> 32) & 0xFFFFFFFF, np.uint32)], axis=-1)
    return philox4x32_10(ctr, key).reshape(-1)[:P]


# --------------------------------------------------------------------------------------
# Helpers shared by tests / bench (synthetic inputs of SURVEY.md 8(d))
# --------------------------------------------------------------------------------------


def split_flat(layers, flat):
    out, o = [], 0
    for s in weight_shapes(layers):
        k = int(np.prod(s))
        out.append(np.asarray(flat[o:o + k]).reshape(s))
        o += k
    return out


def flatten_list(tensors):
    return np.concatenate([np.asarray(t, np.float32).reshape(-1) for t in tensors])


def synthetic_posterior(layers, seed=1, sigma_scale=0.1):
    """mu ~ N(0, 1/fan_in) (biases too), sigma = sigma_scale/sqrt(fan_in) per layer."""
    g = np.random.Generator(np.random.Philox(seed))
    mean, std = [], []
    for l in layers:
        shp = l.weight_shapes()
        if not shp:
            continue
        fan_in = int(np.prod(shp[0][:-1]))
        for s in shp:
            mean.append((g.standard_normal(s) / math.sqrt(fan_in)).astype(np.float32))
            std.append(np.full(s, sigma_scale / math.sqrt(fan_in), np.float32))
    return mean, std


def synthetic_eps(layers, n, seed=2):
    g = np.random.Generator(np.random.Philox(seed))
    return [[g.standard_normal(s).astype(np.float32) for s in weight_shapes(layers)] for _ in range(n)]


def synthetic_x(shape, seed=0):
    g = np.random.Generator(np.random.Philox(seed))
    return g.random(shape, dtype=np.float32)
