"""PosteriorModel -- drop-in for ``deepbayes.PosteriorModel`` (deepbayes/posteriormodel.py)
on the hot path: load a saved posterior, ``sample()``, ``predict(x, n)``, ``predict_logits``,
``_predict``, ``set_weights``.  The Monte-Carlo loop of predict (posteriormodel.py:94-101 --
n x {sample all weights, set_weights, eager forward, += }) is ONE call into the sm_100a
library: weights are drawn on the device (Philox keyed by a global sample id), forwarded
through the fused kernels, and only the [B,C] sample-sums come back.

Reference quirks kept on purpose: ``var.npy`` is used as a standard
deviation (:74-75); ``predict`` leaves the model holding the last sampled weights (:96) while
``sample`` does not touch it (:61-63); ``_predict_logits`` drops the bias and ignores ``n``
(:103-110); a missing ``info.pkl`` falls back to +-inf bounds instead of raising (:56).
"""
from __future__ import annotations

import os
from typing import Optional

import numpy as np

from . import distributed as _dd
from . import formats
from .engine import Engine, flatten_weights
from .keras_compat import Sequential, from_keras


def _split(model: Sequential, flat: np.ndarray):
    out, o = [], 0
    for s in model.weight_shapes():
        k = int(np.prod(s))
        out.append(np.asarray(flat[o:o + k], dtype=np.float32).reshape(s))
        o += k
    return out


class PosteriorModel:
    """:param path_to_model: directory written by any optimizer's ``save``.
    :param deterministic: turn sampling off (posteriormodel.py:32,42-43).
    Extra keyword arguments (not in the reference): ``mode`` = 'bf16' (tcgen05 tensor cores,
    default) or 'fp32' (exact FFMA path, <=1e-5 of the fp32 reference); ``seed`` = Philox key
    (default: drawn from np.random so ``np.random.seed`` controls it, like the reference's
    global RNG); ``device``; ``shard`` = split the n samples over torch.distributed ranks."""

    def __init__(self, path_to_model, deterministic=False, mode="bf16", seed=None, device=None, shard=True, model=None):
        self.path_to_model = path_to_model
        self.mode = mode
        self.shard = shard
        self.model = from_keras(model) if model is not None else formats.load_arch(path_to_model)
        self.seed = int(np.random.randint(0, 2 ** 31 - 1)) if seed is None else int(seed)
        self._next_sample = 0
        self._engine = Engine(self.model, device)
        if formats.is_gaussian_dir(path_to_model):  # posteriormodel.py:36-44
            self.det = deterministic
            self.posterior_mean = formats.load_tensor_list(os.path.join(path_to_model, "mean.npy"))
            self.posterior_var = formats.load_tensor_list(os.path.join(path_to_model, "var.npy"))
            self._engine.set_gaussian(self.posterior_mean, self.posterior_var)
            if self.det:
                self.model.set_weights(self.posterior_mean)
            self.sample_based = False
        else:  # posteriormodel.py:45-54
            self.sample_based = True
            self.det = False
            self.frequency = np.load(os.path.join(path_to_model, "freq.npy"), allow_pickle=True)
            self.frequency = self.frequency / np.sum(self.frequency)
            self.num_post_samps = len(self.frequency)
            self._slab_host = formats.load_slab(path_to_model, self._engine.P, self.num_post_samps)
            self._engine.set_samples(self._slab_host)
            mp = os.path.join(path_to_model, "mean.npy")
            if os.path.exists(mp):
                self.posterior_mean = formats.load_tensor_list(mp)
        info = formats.load_info(path_to_model)  # posteriormodel.py:56-58
        self.input_upper = info["input_upper"]
        self.input_lower = info["input_lower"]

    # ---- sampling ---------------------------------------------------------------------
    def _take_ids(self, n):
        s0 = self._next_sample
        self._next_sample += n
        return s0

    def sample(self, inflate=1.0):
        """posteriormodel.py:60-79.  Returns a fresh list [W0,b0,...] (float64 for Gaussian posteriors, like the reference;
        stored samples come back in their stored dtype); the model is not touched."""
        if self.det:
            return self.model.get_weights()
        if not self.sample_based:
            W = self._engine.sample(1, self.seed, self._take_ids(1), None, inflate)
            # float64 like np.random.normal(loc=..., scale=...) at :74-75 (the values are the device's fp32 draws)
            return [w.astype(np.float64) for w in _split(self.model, W[0].cpu().numpy())]
        index = np.random.choice(range(self.num_post_samps), p=self.frequency)  # :77
        return _split(self.model, self._slab_host[index])

    def set_weights(self, weights):
        self.model.set_weights(weights)

    # ---- prediction -------------------------------------------------------------------
    def _mc_sums(self, x, n, out_kind, second_moment=False, inflate=1.0):
        """sum over n posterior samples of the model output, sharded over ranks if asked."""
        rank, ws = _dd.world() if self.shard else (0, 1)
        if ws > 1 and isinstance(x, np.ndarray) and x.size >= (1 << 18) and _dd.sharded_upload_enabled(ws):
            # every rank holds the same host batch: each uploads 1 / ws of its rows, one all-gather over NVLink completes it
            xg = _dd.sharded_upload(np.ascontiguousarray(x, dtype=np.float32).reshape(-1, self._engine.K), self._engine._dev_f32, self._engine.device)
            x = x if xg is None else xg
        if not self.sample_based:
            s0 = self._take_ids(n)
            off, cnt = _dd.shard_range(n, rank, ws)
            if cnt > 0:
                s1, s2 = self._engine.predict_sums(x, cnt, self.seed, s0 + off, None, inflate, self.mode, out_kind, second_moment)
            else:
                s1, s2 = self._zeros_like_out(x, second_moment)
            last = (self.seed, s0 + n - 1, inflate)
            self.model.set_weights_lazy(lambda: _split(self.model, self._engine.sample(1, last[0], last[1], None, last[2])[0].cpu().numpy()))
        else:
            idx = np.random.choice(self.num_post_samps, size=n, p=self.frequency).astype(np.int32)  # n draws of :77
            off, cnt = _dd.shard_range(n, rank, ws)
            if cnt > 0:
                s1, s2 = self._engine.predict_stored_sums(x, cnt, 0, idx[off:off + cnt], None, self.mode, out_kind, second_moment)
            else:
                s1, s2 = self._zeros_like_out(x, second_moment)
            li = int(idx[-1])
            self.model.set_weights_lazy(lambda: _split(self.model, self._slab_host[li]))
        if ws > 1:
            _dd.allreduce_sum_(s1, s2)
        return s1, s2

    def _zeros_like_out(self, x, second_moment):
        torch = self._engine.torch
        B = int((x.numel() if isinstance(x, torch.Tensor) else np.asarray(x).size) // self._engine.K)
        z = torch.zeros((B, self._engine.C), dtype=torch.float32, device=self._engine.device)
        return z, (torch.zeros_like(z) if second_moment else None)

    def predict(self, input, n=35):
        """Mean of the posterior predictive over n samples (posteriormodel.py:81-101)."""
        if self.det:
            return self.model(input)
        _, oshape = self.model.rows_and_outshape(np.asarray(input))
        s1, _ = self._mc_sums(input, int(n), "probs")
        return (s1 / float(n)).cpu().numpy().reshape(oshape)

    def predict_logits(self, input, n=35):
        """Mean pre-softmax output over n samples (posteriormodel.py:114-139)."""
        _, oshape = self.model.rows_and_outshape(np.asarray(input))
        if self.det:
            return self._engine.forward(input, self.model.get_weights(), "fp32", "logits")
        s1, _ = self._mc_sums(input, int(n), "logits")
        return (s1 / float(n)).cpu().numpy().reshape(oshape)

    def predict_moments(self, input, n=35):
        """(mean, E[p^2]) over n samples in one pass -- what uncertainty.py:15-36 derives its
        epistemic / aleatoric terms from."""
        _, oshape = self.model.rows_and_outshape(np.asarray(input))
        s1, s2 = self._mc_sums(input, int(n), "probs", second_moment=True)
        return (s1 / float(n)).cpu().numpy().reshape(oshape), (s2 / float(n)).cpu().numpy().reshape(oshape)

    def _predict(self, input):
        """Single forward with the current weights (posteriormodel.py:112-113)."""
        return self.model(input)

    def _predict_logits(self, input, n=35):
        """posteriormodel.py:103-110 as written: current weights, NO bias on the last layer,
        ``n`` ignored."""
        ws = self.model.get_weights()
        ws = ws[:-1] + [np.zeros_like(ws[-1])]
        return self._engine.forward(input, ws, "fp32", "logits")
