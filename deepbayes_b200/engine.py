"""Host-side wrapper of one C-ABI handle: owns the device-resident posterior and exposes the
hot-path calls with torch tensors as device memory (torch is plumbing here: allocation,
streams, torch.distributed -- every FLOP is issued by libdeepbayes_b200.so)."""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import numpy as np

from . import _lib


def _torch():
    import torch
    return torch


def _require_cuda(device=None):
    torch = _torch()
    if not torch.cuda.is_available():
        raise _lib.DbxError("deepbayes_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    if device is None:
        device = torch.cuda.current_device()
    return torch.device("cuda", device if isinstance(device, int) else (device.index or 0))


def _ptr(t):
    return C.c_void_p(0 if t is None else t.data_ptr())


def _stream():
    return C.c_void_p(_torch().cuda.current_stream().cuda_stream)


def flatten_weights(ws: Sequence[np.ndarray]) -> np.ndarray:
    return np.concatenate([np.asarray(w, dtype=np.float32).reshape(-1) for w in ws]) if len(ws) else np.zeros(0, np.float32)


def device_activation(name: str, x):
    """``layer.activation(x)`` on the device (verifiers.py:552-555 calls it on small arrays)."""
    torch = _torch()
    dev = _require_cuda()
    lib = _lib.load()
    a = np.asarray(x, dtype=np.float32)
    if a.size == 0:
        return a.copy()
    cols = a.shape[-1] if a.ndim else 1
    t = torch.from_numpy(np.ascontiguousarray(a).reshape(-1, cols)).to(dev)
    _lib.check(lib.dbx_activation(_lib.ACT[name], _ptr(t), t.shape[0], t.shape[1], _stream()))
    return t.cpu().numpy().reshape(a.shape)


class Engine:
    """One model plan + its device-resident posterior."""

    def __init__(self, model, device=None):
        self.lib = _lib.load()
        self.torch = _torch()
        self.device = _require_cuda(device)
        self.model = model
        descs = []
        for l in model.layers:
            d = _lib.LayerDesc()
            d.activation = _lib.ACT[l.activation.name]
            if l.kind == "dense":
                d.kind, d.units = _lib.DENSE, l.units
            elif l.kind == "conv2d":
                d.kind, d.units = _lib.CONV2D, l.filters
                d.kh, d.kw = l.kernel_size
                d.sh, d.sw = l.strides
                d.padding = _lib.PAD[l.padding]
            elif l.kind == "maxpool2d":
                d.kind = _lib.MAXPOOL2D
                d.kh, d.kw = l.pool_size
                d.sh, d.sw = l.strides
            elif l.kind == "flatten":
                d.kind = _lib.FLATTEN
            else:
                raise ValueError(l.kind)
            descs.append(d)
        arr = (_lib.LayerDesc * len(descs))(*descs)
        shp = tuple(int(v) for v in model.feature_shape)
        ishape = (C.c_int32 * len(shp))(*shp)
        h = C.c_void_p()
        _lib.check(self.lib.dbx_create(arr, len(descs), ishape, len(shp), self.device.index or 0, C.byref(h)))
        self.h = h
        v = C.c_int64()
        _lib.check(self.lib.dbx_param_count(h, C.byref(v))); self.P = v.value
        _lib.check(self.lib.dbx_output_dim(h, C.byref(v))); self.C = v.value
        _lib.check(self.lib.dbx_input_dim(h, C.byref(v))); self.K = v.value
        _lib.check(self.lib.dbx_flops_per_forward(h, C.byref(v))); self.flops_per_forward = v.value
        self._slab = None
        self._keep = None

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.lib.dbx_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # ---- helpers ----------------------------------------------------------------------
    def _dev_f32(self, a):
        """fp32 device tensor of `a`.  Host arrays go through dbx_upload_f32: one DMA when the array is page-locked (e.g. the
        .numpy() view of a pinned torch tensor), otherwise pieces through the handle's pinned staging ring, each piece
        guarded by the event of its previous copy -- back-to-back uploads never overwrite data still in flight."""
        torch = self.torch
        if isinstance(a, torch.Tensor):
            if a.is_cuda:
                return a.to(device=self.device, dtype=torch.float32).contiguous()
            a = a.detach().to(torch.float32).contiguous().numpy()
        a = np.ascontiguousarray(np.asarray(a, dtype=np.float32))
        t = torch.empty(a.shape, dtype=torch.float32, device=self.device)
        if a.size:
            _lib.check(self.lib.dbx_upload_f32(self.h, a.ctypes.data_as(C.c_void_p), _ptr(t), a.size, _stream()))
            self._keep = a      # a pinned source must outlive its DMA: keep the latest one referenced
        return t

    def _as_flat(self, a):
        if isinstance(a, self.torch.Tensor):
            return a.reshape(-1)
        if isinstance(a, (list, tuple)) or (isinstance(a, np.ndarray) and a.dtype == object):
            return flatten_weights(list(a))
        return np.asarray(a, dtype=np.float32).reshape(-1)

    def _x2d(self, x):
        t = self._dev_f32(x)
        if t.numel() % self.K != 0:
            raise ValueError("input with %d elements is not a multiple of the model input dim %d" % (t.numel(), self.K))
        return t.reshape(-1, self.K)

    # ---- posterior --------------------------------------------------------------------
    def set_gaussian(self, mu, sigma, rho=False):
        mu_t = self._dev_f32(self._as_flat(mu)).reshape(-1)
        sg_t = self._dev_f32(self._as_flat(sigma)).reshape(-1)
        if mu_t.numel() != self.P or sg_t.numel() != self.P:
            raise ValueError("posterior has %d / %d parameters, model needs %d" % (mu_t.numel(), sg_t.numel(), self.P))
        fn = self.lib.dbx_set_gaussian_rho if rho else self.lib.dbx_set_gaussian
        _lib.check(fn(self.h, _ptr(mu_t), _ptr(sg_t), self.P, 1, _stream()))
        self.torch.cuda.current_stream().synchronize()

    def get_gaussian(self):
        mu = np.empty(self.P, np.float32)
        sg = np.empty(self.P, np.float32)
        _lib.check(self.lib.dbx_get_gaussian(self.h, mu.ctypes.data_as(C.c_void_p), sg.ctypes.data_as(C.c_void_p), self.P, _stream()))
        return mu, sg

    def set_samples(self, slab):
        """slab: [S,P] fp32 (numpy or torch), or [S,pitch] with pitch = P rounded up to 4 (16-byte rows: 128-bit streaming
        loads and TMA without a padded copy -- what a 74.5 GB slab needs); kept alive here, borrowed by the library."""
        t = self._dev_f32(slab)
        pitch = (self.P + 3) // 4 * 4          # 16-byte aligned rows -> 128-bit streaming loads
        if t.ndim != 2 or t.shape[1] not in (self.P, pitch):
            raise ValueError("slab must be [S,%d] (or [S,%d] padded), got %r" % (self.P, pitch, tuple(t.shape)))
        if t.shape[1] != pitch:
            padded = self.torch.zeros((t.shape[0], pitch), dtype=self.torch.float32, device=self.device)
            padded[:, :self.P] = t
            t = padded
        self._slab = t
        self.S = int(t.shape[0])
        _lib.check(self.lib.dbx_set_samples_strided(self.h, _ptr(t), t.shape[0], self.P, pitch))

    # ---- hot path ---------------------------------------------------------------------
    def sample(self, n=1, seed=0, s0=0, eps=None, inflate=1.0, return_eps=False):
        torch = self.torch
        W = torch.empty((n, self.P), dtype=torch.float32, device=self.device)
        e = None if eps is None else self._dev_f32(eps).reshape(n, self.P)
        eo = torch.empty_like(W) if return_eps else None
        _lib.check(self.lib.dbx_sample(self.h, seed, s0, n, _ptr(e), float(inflate), _ptr(W), _ptr(eo), _stream()))
        return (W, eo) if return_eps else W

    def predict_sums(self, x, n, seed=0, s0=0, eps=None, inflate=1.0, mode="bf16", out_kind="probs", second_moment=False,
                     out=None):
        torch = self.torch
        xt = self._x2d(x)
        B = xt.shape[0]
        s1 = out[0] if out is not None else torch.empty((B, self.C), dtype=torch.float32, device=self.device)
        s2 = (out[1] if out is not None else torch.empty_like(s1)) if second_moment else None
        e = None if eps is None else self._dev_f32(eps).reshape(n, self.P)
        _lib.check(self.lib.dbx_predict(self.h, _ptr(xt), B, n, seed, s0, _ptr(e), float(inflate), _lib.MODE[mode],
                                        _lib.OUT_LOGITS if out_kind == "logits" else _lib.OUT_PROBS, 0, _ptr(s1), _ptr(s2),
                                        _stream()))
        return s1, s2

    def predict_stored_sums(self, x, n, j0=0, idx=None, wts=None, mode="fp32", out_kind="probs", second_moment=False):
        torch = self.torch
        xt = self._x2d(x)
        B = xt.shape[0]
        s1 = torch.empty((B, self.C), dtype=torch.float32, device=self.device)
        s2 = torch.empty_like(s1) if second_moment else None
        it = None
        if idx is not None:
            ia = np.asarray(idx, dtype=np.int32).reshape(-1)
            S = getattr(self, "S", 0)
            if ia.size < n or (ia.size and (ia.min() < 0 or ia.max() >= S)):   # the device gathers rows of the slab by these
                raise ValueError("stored-sample indices must be %d values in [0, %d)" % (n, S))
            it = torch.as_tensor(ia).to(self.device)
        elif j0 < 0 or j0 + n > getattr(self, "S", 0):
            raise ValueError("stored-sample range [%d, %d) outside the slab of %d samples" % (j0, j0 + n, getattr(self, "S", 0)))
        wt = None if wts is None else self._dev_f32(wts).reshape(-1)
        _lib.check(self.lib.dbx_predict_stored(self.h, _ptr(xt), B, j0, n, _ptr(it), _ptr(wt), _lib.MODE[mode],
                                               _lib.OUT_LOGITS if out_kind == "logits" else _lib.OUT_PROBS, 0, _ptr(s1),
                                               _ptr(s2), _stream()))
        return s1, s2

    def forward_device(self, x, W, mode="fp32", out_kind="probs"):
        torch = self.torch
        xt = self._x2d(x)
        Wt = self._dev_f32(W).reshape(-1)
        if Wt.numel() != self.P:
            raise ValueError("weight vector has %d elements, model needs %d" % (Wt.numel(), self.P))
        out = torch.empty((xt.shape[0], self.C), dtype=torch.float32, device=self.device)
        _lib.check(self.lib.dbx_forward_one(self.h, _ptr(xt), xt.shape[0], _ptr(Wt), _lib.MODE[mode],
                                            _lib.OUT_LOGITS if out_kind == "logits" else _lib.OUT_PROBS, _ptr(out), _stream()))
        return out

    def forward(self, x, weights, mode="fp32", out_kind="probs"):
        """numpy in / numpy out single forward, output shaped like Keras would."""
        _, oshape = self.model.rows_and_outshape(np.asarray(x))
        out = self.forward_device(np.asarray(x, dtype=np.float32), flatten_weights(weights), mode, out_kind)
        return out.cpu().numpy().reshape(oshape)

    def bbb_forward(self, x, seed=0, step=0, eps=None, mode="fp32", return_weights=False):
        torch = self.torch
        xt = self._x2d(x)
        out = torch.empty((xt.shape[0], self.C), dtype=torch.float32, device=self.device)
        e = None if eps is None else self._dev_f32(eps).reshape(-1)
        W = torch.empty(self.P, dtype=torch.float32, device=self.device) if return_weights else None
        eo = torch.empty(self.P, dtype=torch.float32, device=self.device) if return_weights else None
        _lib.check(self.lib.dbx_bbb_forward(self.h, _ptr(xt), xt.shape[0], seed, step, _ptr(e), _lib.MODE[mode], _ptr(out),
                                            _ptr(W), _ptr(eo), _stream()))
        return (out, W, eo) if return_weights else out

    def count_predicate(self, x, labels, n, seed=0, s0=0, eps=None, inflate=1.0, mode="bf16", return_flags=False):
        torch = self.torch
        xt = self._x2d(x)
        K = xt.shape[0]
        lab = torch.as_tensor(np.asarray(labels, dtype=np.int32).reshape(-1).copy()).to(self.device)
        if lab.numel() != K:
            raise ValueError("need one label per input row (%d rows, %d labels)" % (K, lab.numel()))
        counts = torch.zeros(K, dtype=torch.int64, device=self.device)
        flags = torch.empty((n, K), dtype=torch.uint8, device=self.device) if return_flags else None
        e = None if eps is None else self._dev_f32(eps).reshape(n, self.P)
        _lib.check(self.lib.dbx_count_predicate(self.h, _ptr(xt), K, _ptr(lab), n, seed, s0, _ptr(e), float(inflate),
                                                _lib.MODE[mode], 0, _ptr(flags), _ptr(counts), _stream()))
        return (counts, flags) if return_flags else counts

    def bbb_step(self, x, labels, mean_t, rho_t, prior_mean_t, prior_var_t, lrate, kl_weight=1.0, seed=0, step=0, noise=None,
                 want_probs=True, want_weights=False, robust_mode=0, epsilon=0.0, robust_lambda=1.0, in_lower=-3.0e38, in_upper=3.0e38,
                 eps_draws=None, mode="fp32"):
        """One Bayes-by-Backprop step (bayesbybackprop.py:46-170) updating the flat device tensors ``mean_t`` / ``rho_t`` in
        place.  robust_mode 3 (IBP averaged over the epsilon draws ``eps_draws``, :102-121) goes through `dbx_bbb_step_ibp_mc`.
        mode "bf16" (robust_mode 0 / 2): the step's GEMMs on the tensor cores, parameters and update in fp32.
        Returns dict(loss=[loss, kl] device tensor, probs=[B,C] | None, W=[P] | None)."""
        torch = self.torch
        xt = self._x2d(x)
        B = xt.shape[0]
        lab = torch.as_tensor(np.asarray(labels, dtype=np.int32).reshape(-1).copy()).to(self.device)
        if lab.numel() != B:
            raise ValueError("need one label per input row (%d rows, %d labels)" % (B, lab.numel()))
        for t in (mean_t, rho_t, prior_mean_t, prior_var_t):
            if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float32 and t.numel() == self.P and t.is_contiguous()):
                raise ValueError("parameter vectors must be contiguous fp32 CUDA tensors with %d elements" % self.P)
        loss = torch.empty(2, dtype=torch.float32, device=self.device)
        probs = torch.empty((B, self.C), dtype=torch.float32, device=self.device) if want_probs else None
        W = torch.empty(self.P, dtype=torch.float32, device=self.device) if want_weights else None
        e = None if noise is None else self._dev_f32(noise).reshape(-1)
        if int(robust_mode) == 3:
            import ctypes
            d = np.ascontiguousarray(np.asarray(eps_draws, np.float32).reshape(-1))
            _lib.check(self.lib.dbx_bbb_step_ibp_mc(self.h, _ptr(xt), B, _ptr(lab), _ptr(mean_t), _ptr(rho_t), _ptr(prior_mean_t), _ptr(prior_var_t),
                                                    _ptr(e), seed, step, float(lrate), float(kl_weight), int(d.size),
                                                    d.ctypes.data_as(ctypes.c_void_p), _ptr(loss), _ptr(probs), _ptr(W), _stream()))
            return {"loss": loss, "probs": probs, "W": W}
        _lib.check(self.lib.dbx_bbb_step(self.h, _ptr(xt), B, _ptr(lab), _ptr(mean_t), _ptr(rho_t), _ptr(prior_mean_t), _ptr(prior_var_t),
                                         _ptr(e), seed, step, float(lrate), float(kl_weight), int(robust_mode), float(epsilon),
                                         float(robust_lambda), float(max(in_lower, -3.0e38)), float(min(in_upper, 3.0e38)), _lib.MODE[mode],
                                         _ptr(loss), _ptr(probs), _ptr(W), _stream()))
        return {"loss": loss, "probs": probs, "W": W}

    def bbb_train_epoch(self, x_t, labels_t, batch, mean_t, rho_t, prior_mean_t, prior_var_t, lrate, kl_weight=1.0, seed=0, step0=0,
                        robust_mode=0, epsilon=0.0, robust_lambda=1.0, in_lower=-3.0e38, in_upper=3.0e38, want_weights=True, mode="fp32"):
        """One epoch of Bayes-by-Backprop steps in ONE device call (`dbx_bbb_train_epoch`): ``x_t`` [N,K] fp32 / ``labels_t`` [N]
        int32 CUDA tensors in visiting order.  Returns dict(loss=[nb,2], probs=[N,C], W=[P] | None) device tensors."""
        torch = self.torch
        N = int(x_t.shape[0])
        if not (x_t.is_cuda and x_t.dtype == torch.float32 and x_t.is_contiguous() and x_t.numel() == N * self.K):
            raise ValueError("x_t must be a contiguous fp32 CUDA tensor [N,%d]" % self.K)
        if not (labels_t.is_cuda and labels_t.dtype == torch.int32 and labels_t.is_contiguous() and labels_t.numel() == N):
            raise ValueError("labels_t must be a contiguous int32 CUDA tensor [N]")
        for t in (mean_t, rho_t, prior_mean_t, prior_var_t):
            if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float32 and t.numel() == self.P and t.is_contiguous()):
                raise ValueError("parameter vectors must be contiguous fp32 CUDA tensors with %d elements" % self.P)
        nb = (N + int(batch) - 1) // int(batch)
        loss = torch.empty((nb, 2), dtype=torch.float32, device=self.device)
        probs = torch.empty((N, self.C), dtype=torch.float32, device=self.device)
        W = torch.empty(self.P, dtype=torch.float32, device=self.device) if want_weights else None
        _lib.check(self.lib.dbx_bbb_train_epoch(self.h, _ptr(x_t), _ptr(labels_t), N, int(batch), _ptr(mean_t), _ptr(rho_t), _ptr(prior_mean_t),
                                                _ptr(prior_var_t), seed, step0, float(lrate), float(kl_weight), int(robust_mode), float(epsilon),
                                                float(robust_lambda), float(max(in_lower, -3.0e38)), float(min(in_upper, 3.0e38)), _lib.MODE[mode],
                                                _ptr(loss), _ptr(probs), _ptr(W), _stream()))
        return {"loss": loss, "probs": probs, "W": W}

    def input_gradient(self, x, labels, n_outer=1, n_inner=1, seed=0, s0=0, noise=None, inflate=1.0, stored=False, j0=0, mode="fp32", idx=None):
        """Sum over n_outer iterations of d CE(labels, mean of n_inner sampled predictions) / d x (attacks.py:49-75): [B,K] device tensor.
        mode "bf16": forward and backward GEMMs on the tensor cores (operands rounded to bf16), "fp32": FFMA kernels."""
        torch = self.torch
        xt = self._x2d(x)
        B = xt.shape[0]
        lab = torch.as_tensor(np.asarray(labels, dtype=np.int32).reshape(-1).copy()).to(self.device)
        if lab.numel() != B:
            raise ValueError("need one label per input row (%d rows, %d labels)" % (B, lab.numel()))
        g = torch.empty((B, self.K), dtype=torch.float32, device=self.device)
        e = None if noise is None else self._dev_f32(noise).reshape(n_outer * n_inner, self.P)
        it = None
        if idx is not None:                                   # stored rows drawn by the caller (with replacement, posteriormodel.py:77)
            ia = np.asarray(idx, dtype=np.int64).reshape(-1)
            if not stored or ia.size != n_outer * n_inner or ia.min() < 0 or ia.max() >= getattr(self, "S", 0):
                raise ValueError("idx needs stored=True and n_outer * n_inner rows in [0, %d)" % getattr(self, "S", 0))
            it = torch.as_tensor(ia.astype(np.int32)).to(self.device)
        _lib.check(self.lib.dbx_input_gradient(self.h, _ptr(xt), B, _ptr(lab), n_outer, n_inner, seed, s0, _ptr(e), float(inflate),
                                               1 if stored else 0, j0, _ptr(it), _lib.MODE[mode], _ptr(g), _stream()))
        return g

    def input_gradient_one(self, x, labels, weights_flat):
        torch = self.torch
        xt = self._x2d(x)
        B = xt.shape[0]
        lab = torch.as_tensor(np.asarray(labels, dtype=np.int32).reshape(-1).copy()).to(self.device)
        w = self._dev_f32(weights_flat).reshape(-1)
        g = torch.empty((B, self.K), dtype=torch.float32, device=self.device)
        _lib.check(self.lib.dbx_input_gradient_one(self.h, _ptr(xt), B, _ptr(lab), _ptr(w), _ptr(g), _stream()))
        return g

    def count_predicate_fgsm(self, x, attack_labels, labels, eps, n, seed=0, s0=0, noise=None, inflate=1.0, lower=-3.0e38, upper=3.0e38,
                             return_flags=False, mode="fp32"):
        """Per posterior sample: FGSM against that sample, then argmax == label at the adversarial input (verifiers.py:622-634)."""
        torch = self.torch
        xt = self._x2d(x)
        B = xt.shape[0]
        al = torch.as_tensor(np.asarray(attack_labels, dtype=np.int32).reshape(-1).copy()).to(self.device)
        lab = torch.as_tensor(np.asarray(labels, dtype=np.int32).reshape(-1).copy()).to(self.device)
        if lab.numel() != B or al.numel() != B:
            raise ValueError("need one label per input row")
        counts = torch.zeros(B, dtype=torch.int64, device=self.device)
        flags = torch.empty((n, B), dtype=torch.uint8, device=self.device) if return_flags else None
        e = None if noise is None else self._dev_f32(noise).reshape(n, self.P)
        lo = float(max(lower, -3.0e38)); hi = float(min(upper, 3.0e38))
        _lib.check(self.lib.dbx_count_predicate_fgsm(self.h, _ptr(xt), B, _ptr(al), _ptr(lab), float(eps), lo, hi, n, seed, s0, _ptr(e),
                                                     float(inflate), _lib.MODE[mode], 0, _ptr(flags), _ptr(counts), _stream()))
        return (counts, flags) if return_flags else counts

    def ibp_predict(self, x, eps, labels=None, n=1, seed=0, s0=0, noise=None, inflate=1.0, mode="bf16", clip=True,
                    stored=False, j0=0, want=("worst",), return_flags=False):
        """Interval bound propagation over n posterior samples (verifiers.py:23-41, 68-95, 537-557, 588-634).
        ``want`` picks the outputs: "worst" = sum over samples of softmax(worst-case logits) [B,C], "lower" /
        "upper" = sums of the logit bounds [B,C], "counts" = per input, the number of samples whose worst case
        still has argmax == label (with return_flags also the [n,B] 0/1 flags).  Returns a dict of device tensors."""
        torch = self.torch
        xt = self._x2d(x)
        B = xt.shape[0]
        lab = None
        if labels is not None:
            lab = torch.as_tensor(np.asarray(labels, dtype=np.int32).reshape(-1).copy()).to(self.device)
            if lab.numel() != B:
                raise ValueError("need one label per input row (%d rows, %d labels)" % (B, lab.numel()))
        z = lambda: torch.zeros((B, self.C), dtype=torch.float32, device=self.device)   # noqa: E731
        out = {k: z() for k in ("worst", "lower", "upper") if k in want}
        counts = torch.zeros(B, dtype=torch.int64, device=self.device) if ("counts" in want or return_flags) else None
        flags = torch.empty((n, B), dtype=torch.uint8, device=self.device) if return_flags else None
        e = None if noise is None else self._dev_f32(noise).reshape(n, self.P)
        _lib.check(self.lib.dbx_ibp_predict(self.h, _ptr(xt), B, float(eps), 1 if clip else 0, _ptr(lab), n, seed, s0, _ptr(e),
                                            float(inflate), 1 if stored else 0, j0, _lib.MODE[mode], 0, _ptr(out.get("worst")),
                                            _ptr(out.get("lower")), _ptr(out.get("upper")), _ptr(flags), _ptr(counts), _stream()))
        if counts is not None:
            out["counts"] = counts
        if flags is not None:
            out["flags"] = flags
        return out

    # ---- every sample's own output (callers that apply their own predicate per iteration, verifiers.py:563-653) -------
    def predict_samples(self, x, n, seed=0, s0=0, eps=None, inflate=1.0, mode="fp32", out_kind="probs"):
        """[n,B,C] device tensor: output of posterior sample s0+i at x (what uncertainty.py:15-36 stacks)."""
        torch = self.torch
        xt = self._x2d(x)
        out = torch.empty((n, xt.shape[0], self.C), dtype=torch.float32, device=self.device)
        e = None if eps is None else self._dev_f32(eps).reshape(n, self.P)
        _lib.check(self.lib.dbx_predict_samples(self.h, _ptr(xt), xt.shape[0], n, seed, s0, _ptr(e), float(inflate), _lib.MODE[mode],
                                                _lib.OUT_LOGITS if out_kind == "logits" else _lib.OUT_PROBS, _ptr(out), _stream()))
        return out

    def fgsm_samples(self, x, attack_labels, eps, n, seed=0, s0=0, noise=None, inflate=1.0, lower=-3.0e38, upper=3.0e38, mode="fp32"):
        """[n,B,C] device tensor: softmax output of sample i at ITS OWN FGSM input (verifiers.py:622-625)."""
        torch = self.torch
        xt = self._x2d(x)
        B = xt.shape[0]
        al = torch.as_tensor(np.asarray(attack_labels, dtype=np.int32).reshape(-1).copy()).to(self.device)
        if al.numel() != B:
            raise ValueError("need one attack label per input row")
        out = torch.empty((n, B, self.C), dtype=torch.float32, device=self.device)
        e = None if noise is None else self._dev_f32(noise).reshape(n, self.P)
        lo = float(max(lower, -3.0e38)); hi = float(min(upper, 3.0e38))
        _lib.check(self.lib.dbx_fgsm_samples(self.h, _ptr(xt), B, _ptr(al), float(eps), lo, hi, n, seed, s0, _ptr(e), float(inflate),
                                             _lib.MODE[mode], _ptr(out), _stream()))
        return out

    def ibp_samples(self, x, eps, n, seed=0, s0=0, noise=None, inflate=1.0, mode="fp32", clip=True, box_hi=None):
        """(lower, upper) [n,B,C] device tensors: logit bounds of sample i over the eps-ball around x (verifiers.py:593), or over
        the explicit box [x, box_hi] (the un-clipped box of IBP(..., predict=True), :605-606)."""
        torch = self.torch
        xt = self._x2d(x)
        B = xt.shape[0]
        ht = None if box_hi is None else self._x2d(box_hi)
        lo = torch.empty((n, B, self.C), dtype=torch.float32, device=self.device)
        hi = torch.empty_like(lo)
        e = None if noise is None else self._dev_f32(noise).reshape(n, self.P)
        _lib.check(self.lib.dbx_ibp_samples(self.h, _ptr(xt), _ptr(ht), B, float(eps), 1 if clip else 0, n, seed, s0, _ptr(e),
                                            float(inflate), _lib.MODE[mode], _ptr(lo), _ptr(hi), _stream()))
        return lo, hi

    def ibp_one(self, x, eps, weights_flat, mode="fp32", clip=True):
        """`IBP(model, inp, weights, eps)` with explicit weights (verifiers.py:68-95): (lower, upper) logits [B,C]."""
        torch = self.torch
        xt = self._x2d(x)
        B = xt.shape[0]
        w = self._dev_f32(weights_flat).reshape(-1)
        if w.numel() != self.P:
            raise ValueError("need %d weights, got %d" % (self.P, w.numel()))
        lo = torch.empty((B, self.C), dtype=torch.float32, device=self.device)
        hi = torch.empty_like(lo)
        _lib.check(self.lib.dbx_ibp_one(self.h, _ptr(xt), B, float(eps), 1 if clip else 0, _ptr(w), _lib.MODE[mode], _ptr(lo), _ptr(hi),
                                        _stream()))
        return lo, hi

    def ibp_state(self, lo, hi, weights_flat, weight_margin=0.0):
        """`IBP_state` (verifiers.py:97-123): explicit input box, explicit weights, weight margin in units of sigma."""
        torch = self.torch
        lt, ht = self._x2d(lo), self._x2d(hi)
        B = lt.shape[0]
        w = self._dev_f32(weights_flat).reshape(-1)
        if w.numel() != self.P or ht.shape != lt.shape:
            raise ValueError("need %d weights and matching lower / upper inputs" % self.P)
        lo_o = torch.empty((B, self.C), dtype=torch.float32, device=self.device)
        hi_o = torch.empty_like(lo_o)
        _lib.check(self.lib.dbx_ibp_state(self.h, _ptr(lt), _ptr(ht), B, _ptr(w), float(weight_margin), _ptr(lo_o), _ptr(hi_o), _stream()))
        return lo_o, hi_o

    def predict_host(self, x: np.ndarray, n, seed=0, s0=0, inflate=1.0, mode="bf16", out_kind="probs", variance=False):
        """Host buffers in, host buffers out (`dbx_predict_host`): the H2D copy of x and the D2H copy of the mean happen inside
        the C call on the handle's own stream."""
        a = np.ascontiguousarray(np.asarray(x, dtype=np.float32)).reshape(-1, self.K)
        B = a.shape[0]
        mean = np.empty((B, self.C), np.float32)
        var = np.empty((B, self.C), np.float32) if variance else None
        _lib.check(self.lib.dbx_predict_host(
            self.h, a.ctypes.data_as(C.c_void_p), B, n, seed, s0, float(inflate), _lib.MODE[mode],
            _lib.OUT_LOGITS if out_kind == "logits" else _lib.OUT_PROBS, mean.ctypes.data_as(C.c_void_p),
            None if var is None else var.ctypes.data_as(C.c_void_p)))
        return (mean, var) if variance else mean

    def philox_raw(self, seed, s, P):
        torch = self.torch
        out = torch.empty(P, dtype=torch.int32, device=self.device)
        _lib.check(self.lib.dbx_philox_raw(seed, s, P, _ptr(out), _stream()))
        return out.cpu().numpy().view(np.uint32)

    def last_path(self) -> str:
        return {0: "generic", 1: "fused_tcgen05", 2: "stream", 3: "generic_tc"}[self.lib.dbx_last_path()]
