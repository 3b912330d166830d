"""Run the REFERENCE'S OWN propagate_conv2d (deepbayes/analyzers/verifiers.py:11-21), unmodified from /root/reference,
on NumPy stand-ins for the TensorFlow symbols it uses, and commit what it returns as tests/golden/reference_conv_ibp.npz.

tf.nn.convolution is supplied by this script as a plain NHWC x HWIO, VALID, stride-1 loop (TensorFlow's defaults for that
call); what the fixture pins is the reference's composition -- w_pos / w_neg, which bound meets which half of the kernel,
and the nominal convolution it adds to BOTH bounds -- not TensorFlow's convolution itself.

    python tests/golden/make_reference_conv_ibp_fixtures.py        (needs /root/reference; CPU only)
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


def conv_valid(x, k):
    x, k = np.asarray(x), np.asarray(k)
    kh, kw, cin, cout = k.shape
    n, h, w, c = x.shape
    out = np.zeros((n, h - kh + 1, w - kw + 1, cout), np.result_type(x.dtype, k.dtype))
    for i in range(kh):
        for j in range(kw):
            out += np.einsum("nhwc,co->nhwo", x[:, i:i + out.shape[1], j:j + out.shape[2], :], k[i, j])
    return out


def main():
    tf = tf_numpy_stub.install([])
    tf.maximum, tf.minimum = np.maximum, np.minimum
    import types
    if not hasattr(tf, "nn"):
        tf.nn = types.SimpleNamespace()
    tf.nn.convolution = conv_valid
    sys.path.insert(0, REF)
    from deepbayes.analyzers import verifiers
    g = np.random.Generator(np.random.Philox(77))
    out = {}
    for tag, (kh, kw, cin, cout, h, w) in {"a": (3, 3, 4, 6, 8, 9), "b": (5, 2, 1, 3, 7, 7)}.items():
        W = g.standard_normal((kh, kw, cin, cout)).astype(np.float64) * 0.4
        b = g.standard_normal(cout).astype(np.float64)
        mid = g.random((2, h, w, cin))
        rad = g.random((2, h, w, cin)) * 0.05
        x_l, x_u = mid - rad, mid + rad
        lo, hi = verifiers.propagate_conv2d(W, b, x_l, x_u)
        out.update({tag + "_W": W, tag + "_b": b, tag + "_xl": x_l, tag + "_xu": x_u, tag + "_l": np.asarray(lo), tag + "_u": np.asarray(hi)})
    np.savez_compressed(os.path.join(HERE, "reference_conv_ibp.npz"), **out)
    print("wrote reference_conv_ibp.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
