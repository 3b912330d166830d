"""Run the REFERENCE'S OWN statistical-verification loop (deepbayes/analyzers/verifiers.py: massart_bound_check :563-653 with
okamoto_bound :517-518 and absolute_massart_halting :521-532), unmodified from /root/reference, on the NumPy stand-in for
TensorFlow and commit what it returns -- together with the interval bounds every iteration saw -- as
tests/golden/reference_massart.npz.

Glue around the reference's code (none of it changes what the loop computes):
  * statsmodels is not installed: proportion_confint(count, nobs, alpha, method='beta') is restated from its definition
    (Clopper-Pearson: beta.ppf(alpha/2, count, nobs-count+1), beta.ppf(1-alpha/2, count+1, nobs-count)) on scipy;
  * the loop calls `.numpy()` on IBP's outputs (:594): IBP is wrapped to hand back ndarray subclasses that have it, and the
    wrapper records every iteration's (lower, upper) logits so that a test can replay the same posterior samples;
  * the dtype-preserving tf symbols of make_reference_ibp_fixtures.py.

    python tests/golden/make_reference_massart_fixtures.py        (needs /root/reference; CPU only)
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

SEED = 20260


class _T(np.ndarray):
    def numpy(self):
        return np.asarray(self)


def clopper_pearson(count, nobs, alpha=0.05, method="beta"):
    """statsmodels.stats.proportion.proportion_confint(method='beta')."""
    from scipy.stats import beta
    assert method == "beta"
    lo = beta.ppf(alpha / 2.0, count, nobs - count + 1)
    hi = beta.ppf(1 - alpha / 2.0, count + 1, nobs - count)
    return lo, hi


def main():
    tf = tf_numpy_stub.install([])
    F32 = np.float32
    tf.float64 = np.float64
    tf.cast = lambda a, dtype=None: np.asarray(a, dtype if dtype is not None else F32)
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

    verifiers.proportion_confint = clopper_pearson
    ref_ibp = verifiers.IBP
    log = []

    def ibp_logged(model, inp, weights, eps, predict=False):
        lo, hi = ref_ibp(model, inp, weights, eps, predict=predict)
        log.append((np.asarray(lo, np.float64).copy(), np.asarray(hi, np.float64).copy()))
        return np.asarray(lo).view(_T), np.asarray(hi).view(_T)
    verifiers.IBP = ibp_logged

    out = {}
    path = os.path.join(REF, "Tutorials/PosteriorModels/VOGN_MNIST_Posterior")
    pm = deepbayes.PosteriorModel(path)
    x1 = o.synthetic_x((1, 784), seed=11)
    out["x"] = x1
    pm.model.set_weights(pm.posterior_mean)
    cls_true = int(np.argmax(np.asarray(pm._predict(x1)).reshape(-1)))
    out["okamoto"] = np.asarray([verifiers.okamoto_bound(e, d) for e, d in ((0.1, 0.3), (0.05, 0.1), (0.25, 0.3))])
    out["halting_args"] = np.asarray([[30, 40, 0.6, 0.9, 0.1, 0.3, 0.05], [3, 40, 0.02, 0.2, 0.1, 0.3, 0.05], [20, 40, 0.35, 0.65, 0.1, 0.3, 0.05],
                                      [400, 400, 0.99, 1.0, 0.05, 0.1, 0.05]], np.float64)
    out["halting"] = np.asarray([verifiers.absolute_massart_halting(int(a[0]), int(a[1]), [a[2], a[3]], a[4], a[5], a[6]) for a in out["halting_args"]])
    wide = [np.asarray(v) * 150.0 for v in pm.posterior_var]          # a posterior wide enough for mixed outcomes
    cases = {  # name: (posterior_var, eps, cls, confidence, delta)
        "all_pass": (pm.posterior_var, 0.001, cls_true, 0.9, 0.3),
        "all_fail": (pm.posterior_var, 0.001, (cls_true + 1) % 10, 0.9, 0.3),
        "mixed": (wide, 0.01, cls_true, 0.85, 0.3),
    }
    for name, (var, eps, cls, conf, delta) in cases.items():
        pm.posterior_var = var
        del log[:]
        np.random.seed(SEED)
        outcomes = []

        def predicate(worst, _cls=cls):
            ok = bool(np.argmax(np.squeeze(worst)) == _cls)
            outcomes.append(ok)
            return ok
        p, it, _mean = verifiers.massart_bound_check(pm, x1, eps, predicate, cls=cls, classification=True, verify=True, confidence=conf, delta=delta)
        out[name + "_params"] = np.asarray([eps, cls, conf, delta], np.float64)
        out[name + "_result"] = np.asarray([p, it], np.float64)
        out[name + "_lo"] = np.stack([l.reshape(-1) for l, _ in log])
        out[name + "_hi"] = np.stack([h.reshape(-1) for _, h in log])
        out[name + "_outcomes"] = np.asarray(outcomes, np.uint8)
        print(name, "p = %.4f after %d iterations (%d IBP calls, %d successes)" % (p, it, len(log), sum(outcomes)))
    np.savez_compressed(os.path.join(HERE, "reference_massart.npz"), **out)
    print("wrote reference_massart.npz:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
