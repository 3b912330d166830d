"""Run the REFERENCE'S OWN zeroth-order attack source (deepbayes/analyzers/attacks.py: zeroth_order_gradient :17-45 and
FGSM :81-102 with order=0), unmodified from /root/reference, on the NumPy stand-in for TensorFlow and commit what it returns as
tests/golden/reference_attacks.npz.  These two functions touch TensorFlow only through tf.eye /
tf.reshape and call the model through model.predict, so the stand-in executes them exactly; the first-order functions
(gradient_expectation, PGD, CW) take their gradient from tf.GradientTape, which the stand-in cannot run -- they stay pinned on
finite differences of the restated loss (tests/test_oracle.py).

    python tests/golden/make_reference_attack_fixtures.py        (needs /root/reference; CPU only)
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


def keras_sparse_ce(y, p):
    """tf.keras.losses.sparse_categorical_crossentropy on ONE prediction row (how zeroth_order_gradient calls it, :36, :39)."""
    p = np.asarray(p, np.float32).reshape(-1)
    return np.float32(-np.log(np.clip(p[int(y)], np.float32(1e-7), np.float32(1 - 1e-7))))


class FixtureModel:
    """The object protocol the attacks rely on (attacks.py:2-7): predict(x, n=...), input_lower / input_upper.  predict = one
    forward of the oracle's Dense MLP with fixed weights (an optimizer object's predict, optimizer.py:293-298)."""

    def __init__(self, oracle, layers, weights, lower=0.0, upper=1.0):
        self.o, self.L, self.W = oracle, layers, weights
        self.input_lower, self.input_upper, self.det = lower, upper, True

    def predict(self, x, n=1):
        x = np.asarray(x, np.float32)
        lead = x.shape[:-1]
        p = self.o.forward(self.L, self.W, x.reshape(-1, x.shape[-1]), np.float32)
        return p.reshape(lead + (p.shape[-1],))


def main():
    tf = tf_numpy_stub.install([])
    tf.eye = lambda n: np.eye(int(n), dtype=np.float32)
    tf.reshape = lambda a, shape: np.reshape(np.asarray(a), shape)
    sys.path.insert(0, REF)
    from deepbayes.analyzers import attacks
    from oracle import oracle as o

    dims = [24, 32, 5]
    L = o.mlp(dims, hidden_act="tanh")
    mean, _ = o.synthetic_posterior(L, seed=3, sigma_scale=1.0)
    W = [np.asarray(w, np.float32) * np.float32(3.0) for w in mean]
    model = FixtureModel(o, L, W)
    out = {"W_flat": o.flatten_list(W).astype(np.float32), "dims": np.asarray(dims)}
    x = o.synthetic_x((3, 24), seed=4)
    lab = np.array([1, 4, 0])
    out["x"], out["labels"], out["step"] = x, lab, np.asarray([0.05], np.float32)
    out["zog"] = np.asarray(attacks.zeroth_order_gradient(model, x, lab, keras_sparse_ce, num_models=1, step=0.05), np.float32)
    # rank-3 inputs like the tutorials' (B, 1, 784)
    x3 = o.synthetic_x((2, 1, 24), seed=6)
    out["x3"] = x3
    out["zog3"] = np.asarray(attacks.zeroth_order_gradient(model, x3, np.array([2, 3]), keras_sparse_ce, num_models=1, step=0.05), np.float32)
    # FGSM with the zeroth-order gradient, direction = -1 (the model's own prediction, :86-91); default step 0.35
    out["fgsm_eps"] = np.asarray([0.1], np.float32)
    out["fgsm0_adv"] = np.asarray(attacks.FGSM(model, x, keras_sparse_ce, 0.1, direction=-1, order=0, samples=1), np.float32)
    np.savez_compressed(os.path.join(HERE, "reference_attacks.npz"), **out)
    print("wrote reference_attacks.npz:", {k: np.asarray(v).shape for k, v in out.items()})


if __name__ == "__main__":
    main()
