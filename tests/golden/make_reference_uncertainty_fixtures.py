"""Run the REFERENCE'S OWN uncertainty source (deepbayes/analyzers/uncertainty.py: variational_uncertainty :15-36,
predictive_entropy :42-51, likelihood_ratio :57-71 -- plain NumPy on top of model.predict), unmodified from /root/reference,
and commit what it returns as tests/golden/reference_uncertainty.npz.  The model object is the
oracle's Monte-Carlo predict over a Gaussian posterior with the Philox sample ids a PosteriorModel would take.

    python tests/golden/make_reference_uncertainty_fixtures.py        (needs /root/reference; CPU only)
"""
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)


class McModel:
    """predict(x, n) = mean over the next n posterior samples (posteriormodel.py:81-101) with Philox normals keyed by the running
    sample id; predict_moments(x, n) = (E[p], E[p^2]) over the next n samples -- what the drop-in PosteriorModel offers."""
    det = False

    def __init__(self, oracle, layers, mean, std, seed=7):
        self.o, self.L, self.mean, self.std, self.seed, self.next = oracle, layers, mean, std, seed, 0

    def _samples(self, x, n):
        o = self.o
        P = o.param_count(self.L)
        out = []
        for s in range(self.next, self.next + n):
            w = o.sample_gaussian(self.mean, self.std, o.split_flat(self.L, o.philox_normals(self.seed, s, P)), order="f32")
            out.append(o.forward(self.L, w, np.asarray(x, np.float32), np.float32))
        self.next += n
        return np.asarray(out)

    def predict(self, x, n=35):
        return self._samples(x, n).mean(0, dtype=np.float32)

    def predict_moments(self, x, n=35):
        p = self._samples(x, n).astype(np.float64)
        return p.mean(0).astype(np.float32), (p ** 2).mean(0).astype(np.float32)


def main():
    spec = importlib.util.spec_from_file_location("ref_uncertainty", os.path.join(REF, "deepbayes/analyzers/uncertainty.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    from oracle import oracle as o
    L = o.mlp([20, 16, 6])
    mean, std = o.synthetic_posterior(L, seed=5, sigma_scale=4.0)
    x, x_out = o.synthetic_x((7, 20), seed=8), o.synthetic_x((9, 20), seed=9) * 3.0
    out = {"x": x, "x_out": x_out, "n": np.asarray([12])}
    m = McModel(o, L, mean, std)
    epi, ale = ref.variational_uncertainty(m, x, num_samples=12)
    out["epistemic"], out["aleatoric"] = np.asarray(epi), np.asarray(ale)
    m = McModel(o, L, mean, std)
    out["entropy"] = np.asarray(ref.predictive_entropy(m, x, num_models=12))
    m = McModel(o, L, mean, std)
    out["likelihood_ratio"] = np.asarray([ref.likelihood_ratio(m, x, x_out, num_models=12)])
    np.savez_compressed(os.path.join(HERE, "reference_uncertainty.npz"), **out)
    print("wrote reference_uncertainty.npz:", {k: np.asarray(v).shape for k, v in out.items()})


if __name__ == "__main__":
    main()
