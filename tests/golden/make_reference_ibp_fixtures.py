"""Run the REFERENCE'S OWN interval-bound-propagation source (deepbayes/analyzers/verifiers.py:
propagate_interval :23-41, IBP :68-95, chernoff_bound_verification :537-557), unmodified from
/root/reference, on the NumPy stand-in for TensorFlow and commit what it returns as
tests/golden/reference_ibp.npz.

On top of tf_numpy_stub.install() this script supplies the dtype-PRESERVING elementwise symbols
the IBP code needs (tf.cast honouring float64, add/subtract/divide/abs/clip_by_value/one_hot):
propagate_interval works in float64 and the fp32-casting versions the other fixture script uses
would silently change that.

    python tests/golden/make_reference_ibp_fixtures.py        (needs /root/reference; CPU only)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import tf_numpy_stub  # noqa: E402

SEED = 424242


def main():
    tf = tf_numpy_stub.install([])
    F32 = np.float32
    tf.float64 = np.float64
    tf.cast = lambda a, dtype=None: np.asarray(a, dtype if dtype is not None else F32)
    # Python scalars stay Python scalars (TensorFlow converts them to the tensor operand's dtype; np.asarray(0.02)
    # would be a float64 0-d array and promote an fp32 input box to float64)
    arr = lambda v: v if isinstance(v, (int, float)) else np.asarray(v)   # noqa: E731
    tf.divide = lambda a, b: arr(a) / arr(b)
    tf.add = lambda a, b: arr(a) + arr(b)
    tf.subtract = lambda a, b: arr(a) - arr(b)
    tf.abs = np.abs
    tf.matmul = lambda a, b: np.asarray(a) @ np.asarray(b)
    tf.clip_by_value = lambda a, lo, hi: np.clip(a, lo, hi)
    tf.one_hot = lambda i, depth: np.eye(depth, dtype=F32)[np.asarray(i)]
    tf.squeeze = np.squeeze
    tf.maximum, tf.minimum = np.maximum, np.minimum
    tf.math.add = tf.add
    tf.math.subtract = tf.subtract
    tf.math.multiply = lambda a, b: arr(a) * arr(b)
    tf.math.abs = np.abs
    tf.math.divide = tf.divide
    sys.path.insert(0, REF)
    import deepbayes
    from deepbayes.analyzers import verifiers
    from oracle import oracle as o

    out = {}
    path = os.path.join(REF, "Tutorials/PosteriorModels/VOGN_MNIST_Posterior")
    pm = deepbayes.PosteriorModel(path)
    x = o.synthetic_x((4, 784), seed=11)                     # fp32, MNIST-normalised range
    eps = 0.02
    out["x"], out["eps"] = x, np.asarray([eps])
    # ---- IBP with explicit weights (posterior mean, then one sampled set)
    pm.model.set_weights(pm.posterior_mean)
    lo, hi = verifiers.IBP(pm, x, pm.model.get_weights(), eps, predict=False)
    out["ibp_mean_l"], out["ibp_mean_u"] = np.asarray(lo), np.asarray(hi)
    lo, hi = verifiers.IBP(pm, x, pm.model.get_weights(), eps, predict=True)
    out["ibp_mean_predict_l"], out["ibp_mean_predict_u"] = np.asarray(lo), np.asarray(hi)
    np.random.seed(SEED)
    w = pm.sample()
    pm.set_weights(w)
    lo, hi = verifiers.IBP(pm, x, pm.model.get_weights(), eps, predict=False)
    out["ibp_sample_l"], out["ibp_sample_u"] = np.asarray(lo), np.asarray(hi)
    out["ibp_sample_W1"] = np.asarray(pm.model.get_weights()[2])
    # ---- chernoff_bound_verification on ONE input (the reference's one_hot(cls) form), 16 samples
    cls = 3
    x1 = x[:1]
    np.random.seed(SEED)
    total = verifiers.chernoff_bound_verification(pm, x1, eps, cls, confidence=0.75, delta=0.3)
    out["chernoff_sum"], out["chernoff_cls"] = np.asarray(total), np.asarray([cls])
    out["chernoff_n"] = np.asarray([int(np.ceil((1 / (2 * 0.25 ** 2)) * np.log(2 / 0.3)))])
    # ---- IBP_state: explicit input box, weight margin 0.5 sigma around the posterior mean (verifiers.py:97-123)
    pm.model.set_weights(pm.posterior_mean)
    s0, s1 = np.clip(x - 0.01, 0, 1).astype(np.float32), np.clip(x + 0.01, 0, 1).astype(np.float32)
    lo, hi = verifiers.IBP_state(pm, s0, s1, pm.model.get_weights(), weight_margin=0.5)
    out["state_s0"], out["state_s1"], out["state_l"], out["state_u"] = s0, s1, np.asarray(lo), np.asarray(hi)
    np.savez_compressed(os.path.join(HERE, "reference_ibp.npz"), **out)
    print("wrote reference_ibp.npz:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
