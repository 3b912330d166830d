"""Regenerate tests/golden/oracle_cases.npz and the fixture copies of the reference's two
saved posteriors.  Run from the repo root:  python tests/golden/make_golden.py
(the reference checkout is only needed for the posterior copies)."""
import os
import shutil
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import cases  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

if __name__ == "__main__":
    blob = {}
    for name in cases.NAMES:
        for k, v in cases.expected(name).items():
            blob["%s/%s" % (name, k)] = v
    np.savez_compressed(os.path.join(HERE, "oracle_cases.npz"), **blob)
    print("wrote oracle_cases.npz with %d arrays" % len(blob))
    ref = "/root/reference/Tutorials/PosteriorModels"
    if os.path.isdir(ref):
        for d in ("VOGN_MNIST_Posterior", "Robust_MNIST_Posterior"):
            dst = os.path.join(HERE, "posteriors", d)
            os.makedirs(dst, exist_ok=True)
            for f in ("mean.npy", "var.npy", "arch.json", "info.pkl"):  # data fixtures, not sources
                shutil.copyfile(os.path.join(ref, d, f), os.path.join(dst, f))
        print("copied the two saved posteriors (without model.h5)")
