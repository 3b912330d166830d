"""Run the REFERENCE'S OWN source (unmodified, from /root/reference) on top of the NumPy stand-in
for TensorFlow (tf_numpy_stub.py) with seeded RNGs, and commit what it returns as
tests/golden/reference_run.npz.  tests/test_oracle.py then checks the oracle against these
outputs (CPU) and tests/test_gpu_parity.py checks the CUDA path against the same vectors.

    python tests/golden/make_reference_fixtures.py        (needs /root/reference; CPU only)
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

SEED_PM, SEED_BBB = 20240601, 777


def main():
    noise_log = []
    tf_numpy_stub.install(noise_log)
    sys.path.insert(0, REF)
    import deepbayes                                     # -> posteriormodel.py (reference source)
    from deepbayes.optimizers import bayesbybackprop     # -> optimizer.py, losses.py, analyzers/*
    from deepbayes.optimizers import losses
    from oracle import oracle as o

    out = {}
    # ---- 1. PosteriorModel on the reference's saved VOGN posterior
    path = os.path.join(REF, "Tutorials/PosteriorModels/VOGN_MNIST_Posterior")
    x = o.synthetic_x((6, 1, 784), seed=0).astype(np.float64)      # tutorials pass float64 (B,1,784)
    pm = deepbayes.PosteriorModel(path)
    np.random.seed(SEED_PM)
    out["pm_predict_n7"] = np.asarray(pm.predict(x, n=7))
    out["pm_weights_after_predict_W1"] = np.asarray(pm.model.get_weights()[2])
    np.random.seed(SEED_PM)
    out["pm_predict_logits_n5"] = np.asarray(pm.predict_logits(x, n=5))
    np.random.seed(SEED_PM)
    s = pm.sample(inflate=2.0)
    out["pm_sample_inflate2_b0"] = np.asarray(s[1])
    out["pm_sample_dtype_is_f64"] = np.asarray([s[0].dtype == np.float64])
    pd = deepbayes.PosteriorModel(path, deterministic=True)
    out["pm_det_predict"] = np.asarray(pd.predict(x, n=50))
    pm.model.set_weights(pm.posterior_mean)
    out["pm__predict_logits_nobias"] = np.asarray(pm._predict_logits(x))

    # ---- 2. BayesByBackprop compile / step (sampling + forward) / sample
    Sequential, Dense = tf_numpy_stub.Sequential, tf_numpy_stub.Dense
    model = Sequential([Dense(24, "relu", input_shape=(30,)), Dense(10, "softmax")])
    captured = []
    metric = lambda labels, preds: captured.append(np.asarray(preds))  # noqa: E731
    losses.KL_Loss = lambda *a, **k: (np.float32(0.0), np.float32(0.0))   # training objective: out of scope
    opt = bayesbybackprop.BayesByBackprop().compile(model, loss_fn=None, batch_size=16, inflate_prior=2.0, metric=metric)
    out["bbb_epochs_after_compile"] = np.asarray([opt.epochs])
    for i, (m, v, r) in enumerate(zip(opt.prior_mean, opt.prior_var, opt.posterior_var)):
        out["bbb_prior_mean_%d" % i], out["bbb_prior_var_%d" % i], out["bbb_rho_%d" % i] = np.asarray(m), np.asarray(v), np.asarray(r)
    xb = o.synthetic_x((16, 30), seed=3)
    np.random.seed(SEED_BBB)
    del noise_log[:]
    opt.step(xb, np.zeros(16, np.int64), 0.1)
    out["bbb_step_predictions"] = captured[-1]
    for i, z in enumerate(noise_log):
        out["bbb_step_noise_%d" % i] = z
    out["bbb_step_model_W0"] = np.asarray(model.get_weights()[0])
    np.random.seed(SEED_BBB)
    sm = opt.sample()
    out["bbb_sample_W0"] = np.asarray(sm[0], np.float64)
    np.savez_compressed(os.path.join(HERE, "reference_run.npz"), **out)
    print("wrote reference_run.npz:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
