"""On-disk posterior formats of the reference.

Gaussian posterior   <dir>/{mean.npy, var.npy, arch.json[, info.pkl][, model.h5]}
                     (optimizer.py:259-276, bayesbybackprop.py:184-195, blrvi.py:171-180)
sample posterior     <dir>/{mean.npy, freq.npy, samples/sample_<i>.npy, arch.json}
                     (hmc.py:308-322, sghmc.py:313-327)
``mean.npy``/``var.npy``/``sample_i.npy`` are pickled object arrays of per-tensor arrays; files
written by the real deepbayes pickle TF EagerTensors, whose only TF symbol is
``tensorflow.python.framework.ops.convert_to_tensor`` -- a 4-line in-memory stub lets NumPy
unpickle them without TensorFlow.  ``model.h5`` is never read (arch.json + mean.npy carry
the same information) and never written (no h5py on the image).
Extra, ours: ``slab.npy`` = the stored samples packed as one [S,P] fp32 matrix for direct
upload to HBM (removes the per-draw ``np.load`` of posteriormodel.py:78).
"""
from __future__ import annotations

import contextlib
import json
import math
import os
import pickle
import sys
import types
from typing import List, Optional

import numpy as np

_TF_MODS = ["tensorflow", "tensorflow.python", "tensorflow.python.framework", "tensorflow.python.framework.ops"]


@contextlib.contextmanager
def tf_unpickle_stub():
    """Temporarily provide ``tensorflow.python.framework.ops.convert_to_tensor`` -> np.asarray
    when TensorFlow is absent, so pickles of EagerTensors load as ndarrays."""
    installed = []
    try:
        if "tensorflow" not in sys.modules:
            try:
                import tensorflow  # noqa: F401  (real TF present: nothing to do)
            except Exception:
                for name in _TF_MODS:
                    sys.modules[name] = types.ModuleType(name)
                    installed.append(name)
                sys.modules[_TF_MODS[-1]].convert_to_tensor = lambda x, *a, **k: np.asarray(x)
        yield
    finally:
        for name in installed:
            sys.modules.pop(name, None)


def load_tensor_list(path: str) -> List[np.ndarray]:
    with tf_unpickle_stub():
        arr = np.load(path, allow_pickle=True)
    return [np.asarray(a, dtype=np.float32) for a in arr]


def save_tensor_list(path: str, tensors) -> None:
    """np.save(path, np.asarray(list_of_arrays)) as the reference does (optimizer.py:262-263),
    always as an object array so ragged shapes are kept."""
    obj = np.empty(len(tensors), dtype=object)
    for i, t in enumerate(tensors):
        obj[i] = np.asarray(t, dtype=np.float32)
    np.save(path, obj, allow_pickle=True)


def load_info(path: str) -> dict:
    """info.pkl (optimizer.py:265-271).  BBB/HMC/SGHMC/SWAG never write it although
    PosteriorModel reads it (posteriormodel.py:56): default to the bounds base compile sets
    (optimizer.py:91-92)."""
    f = os.path.join(path, "info.pkl")
    if os.path.exists(f):
        with open(f, "rb") as fh:
            d = pickle.load(fh)
        return dict(d)
    return {"input_upper": math.inf, "input_lower": -math.inf}


def save_info(path: str, info: dict) -> None:
    with open(os.path.join(path, "info.pkl"), "wb") as f:
        pickle.dump(dict(info), f, pickle.HIGHEST_PROTOCOL)


def load_arch(path: str):
    from .keras_compat import model_from_json
    with open(os.path.join(path, "arch.json")) as f:
        return model_from_json(f.read())


def is_gaussian_dir(path: str) -> bool:
    return os.path.exists(os.path.join(path, "var.npy"))  # posteriormodel.py:36


def count_samples(path: str) -> int:
    d = os.path.join(path, "samples")
    n = 0
    while os.path.exists(os.path.join(d, "sample_%s.npy" % n)):
        n += 1
    return n


def load_slab(path: str, P: int, S: Optional[int] = None) -> np.ndarray:
    """All stored samples as one [S,P] fp32 matrix (slab.npy if present, else packed from the
    per-sample pickles of hmc.py:316-317)."""
    f = os.path.join(path, "slab.npy")
    if os.path.exists(f):
        slab = np.load(f, mmap_mode="r")
        if slab.ndim != 2 or slab.shape[1] != P:
            raise ValueError("slab.npy has shape %r, expected [S,%d]" % (slab.shape, P))
        if S is not None and slab.shape[0] != S:   # a stale / short slab would be indexed out of bounds on the device
            raise ValueError("slab.npy holds %d samples but freq.npy lists %d: delete the stale slab.npy" % (slab.shape[0], S))
        return np.ascontiguousarray(slab, dtype=np.float32)
    S = count_samples(path) if S is None else S
    if S == 0:
        raise FileNotFoundError("no samples/sample_<i>.npy under %s" % path)
    slab = np.empty((S, P), np.float32)
    for i in range(S):
        ts = load_tensor_list(os.path.join(path, "samples", "sample_%s.npy" % i))
        flat = np.concatenate([t.reshape(-1) for t in ts])
        if flat.size != P:
            raise ValueError("sample_%d has %d parameters, model needs %d" % (i, flat.size, P))
        slab[i] = flat
    return slab


def save_sample_posterior(path: str, model, mean, samples, freq, write_slab=True) -> None:
    """hmc.py:308-322 (the leading zero of num_rets is stripped by the caller, :309-310)."""
    os.makedirs(os.path.join(path, "samples"), exist_ok=True)
    save_tensor_list(os.path.join(path, "mean"), mean)
    for i, s in enumerate(samples):
        save_tensor_list(os.path.join(path, "samples", "sample_%s" % i), s)
    np.save(os.path.join(path, "freq"), np.asarray(freq))
    model.save(os.path.join(path, "model.h5"))
    with open(os.path.join(path, "arch.json"), "w") as f:
        f.write(model.to_json())
    if write_slab:
        slab = np.stack([np.concatenate([np.asarray(t, np.float32).reshape(-1) for t in s]) for s in samples])
        np.save(os.path.join(path, "slab.npy"), slab.astype(np.float32))


def save_gaussian_posterior(path: str, model, mean, var, info: Optional[dict] = None) -> None:
    """optimizer.py:259-276 / bayesbybackprop.py:184-195."""
    os.makedirs(path, exist_ok=True)
    save_tensor_list(os.path.join(path, "mean"), mean)
    save_tensor_list(os.path.join(path, "var"), var)
    if info is not None:
        save_info(path, info)
    model.save(os.path.join(path, "model.h5"))
    with open(os.path.join(path, "arch.json"), "w") as f:
        f.write(model.to_json())
