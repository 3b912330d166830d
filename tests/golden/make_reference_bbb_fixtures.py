"""Run the REFERENCE'S OWN loss code (deepbayes/optimizers/losses.py: log_q :10-16, KL_Loss :40-45), unmodified from
/root/reference, on the NumPy stand-in for TensorFlow with a REAL Normal.log_prob, and commit what it returns as
tests/golden/reference_bbb.npz.  The gradients of BayesByBackprop.step come from
tf.GradientTape, which the stand-in cannot provide: the oracle's hand-written gradients are checked against finite
differences of this (pinned) loss instead (tests/test_oracle.py).

    python tests/golden/make_reference_bbb_fixtures.py        (needs /root/reference; CPU only)
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import tf_numpy_stub  # noqa: E402


def main():
    tf = tf_numpy_stub.install([])
    F32 = np.float32
    tf.math.reduce_mean = lambda a, axis=None: np.mean(np.asarray(a), axis=axis)
    tf.cast = lambda a, dtype=None: np.asarray(a, F32 if dtype in (None, 'float32') else dtype)
    tfp = sys.modules["tensorflow_probability"]

    class Normal:   # tfp.distributions.Normal(loc, scale).log_prob, fp32 like TensorFlow's default
        def __init__(self, loc, scale):
            self.loc, self.scale = np.asarray(loc, F32), np.asarray(scale, F32)

        def log_prob(self, x):
            z = (np.asarray(x, F32) - self.loc) / self.scale
            return (F32(-0.5) * z * z - np.log(self.scale) - F32(0.5 * np.log(2.0 * np.pi))).astype(F32)
    tfp.distributions.Normal = Normal
    sys.path.insert(0, REF)
    from deepbayes.optimizers import losses
    from oracle import oracle as o

    L = o.mlp([20, 16, 12, 5])
    mean, std = o.synthetic_posterior(L, seed=3, sigma_scale=2.0)
    rho = [o.inverse_softplus(s) for s in std]
    pm, pv = o.implicit_prior(L, 2.0)
    noise = o.synthetic_eps(L, 1, seed=4)[0]
    W = o.sample_gaussian(mean, std, noise, order="f32")
    x = o.synthetic_x((9, 20), seed=5)
    labels = np.arange(9) % 5
    pred = o.forward(L, W, x)
    sce = lambda y, p: o.sparse_categorical_crossentropy(y, p, np.float32)   # noqa: E731  (the user's loss_fn)
    out = {"labels": labels, "pred": pred}
    for kw in (1.0, 0.25):
        loss, klc = losses.KL_Loss(labels, pred, W, pm, pv, mean, rho, sce, kw)
        out["loss_kw%g" % kw], out["klc_kw%g" % kw] = np.asarray(loss, np.float64), np.asarray(klc, np.float64)
    out["logq_post"] = np.asarray(losses.log_q(W, mean, rho), np.float64)
    out["logq_prior"] = np.asarray(losses.log_q(W, pm, pv), np.float64)
    np.savez_compressed(os.path.join(HERE, "reference_bbb.npz"), **out)
    print("wrote reference_bbb.npz:", {k: np.asarray(v).tolist() if np.asarray(v).size < 3 else np.asarray(v).shape for k, v in out.items()})


if __name__ == "__main__":
    main()
