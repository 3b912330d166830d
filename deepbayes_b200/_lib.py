"""ctypes binding of libdeepbayes_b200.so (the C ABI of include/deepbayes_b200.h).

There is no CPU fallback: if the library is missing or no sm_100 device is present the
compute entry points raise.  Importing this module never touches CUDA.
"""
from __future__ import annotations

import ctypes as C
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libdeepbayes_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "deepbayes_b200.h")

DENSE, CONV2D, MAXPOOL2D, FLATTEN = 0, 1, 2, 3
ACT = {"linear": 0, None: 0, "relu": 1, "softmax": 2, "tanh": 3, "sigmoid": 4}
PAD = {"valid": 0, "same": 1}
MODE = {"fp32": 0, "bf16": 1}
OUT_PROBS, OUT_LOGITS = 0, 1


class LayerDesc(C.Structure):
    _fields_ = [(k, C.c_int32) for k in ("kind", "activation", "units", "kh", "kw", "sh", "sw", "padding")]


class DbxError(RuntimeError):
    pass


_lib = None

_p = C.c_void_p
_i32, _i64, _u64, _f = C.c_int32, C.c_int64, C.c_uint64, C.c_float
_SIGS = {
    "dbx_last_error": (C.c_char_p, []),
    "dbx_version": (C.c_int, []),
    "dbx_launch_count": (_i64, []),
    "dbx_last_path": (C.c_int, []),
    "dbx_profile": (C.c_int, [C.c_int]),
    "dbx_profile_read": (C.c_int, [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(_i64)]),
    "dbx_debug_trace": (C.c_int, [_p, C.c_int]),
    "dbx_set_option": (C.c_int, [C.c_char_p, _i64]),
    "dbx_get_option": (C.c_int, [C.c_char_p, C.POINTER(_i64)]),
    "dbx_reset_options": (C.c_int, []),
    "dbx_create": (C.c_int, [C.POINTER(LayerDesc), _i32, C.POINTER(_i32), _i32, _i32, C.POINTER(_p)]),
    "dbx_destroy": (C.c_int, [_p]),
    "dbx_param_count": (C.c_int, [_p, C.POINTER(_i64)]),
    "dbx_output_dim": (C.c_int, [_p, C.POINTER(_i64)]),
    "dbx_input_dim": (C.c_int, [_p, C.POINTER(_i64)]),
    "dbx_flops_per_forward": (C.c_int, [_p, C.POINTER(_i64)]),
    "dbx_set_gaussian": (C.c_int, [_p, _p, _p, _i64, _i32, _p]),
    "dbx_set_gaussian_rho": (C.c_int, [_p, _p, _p, _i64, _i32, _p]),
    "dbx_get_gaussian": (C.c_int, [_p, _p, _p, _i64, _p]),
    "dbx_set_samples": (C.c_int, [_p, _p, _i64, _i64]),
    "dbx_set_samples_strided": (C.c_int, [_p, _p, _i64, _i64, _i64]),
    "dbx_sample": (C.c_int, [_p, _u64, _i64, _i64, _p, _f, _p, _p, _p]),
    "dbx_predict": (C.c_int, [_p, _p, _i64, _i64, _u64, _i64, _p, _f, _i32, _i32, _i32, _p, _p, _p]),
    "dbx_predict_stored": (C.c_int, [_p, _p, _i64, _i64, _i64, _p, _p, _i32, _i32, _i32, _p, _p, _p]),
    "dbx_forward_one": (C.c_int, [_p, _p, _i64, _p, _i32, _i32, _p, _p]),
    "dbx_bbb_forward": (C.c_int, [_p, _p, _i64, _u64, _i64, _p, _i32, _p, _p, _p, _p]),
    "dbx_count_predicate": (C.c_int, [_p, _p, _i64, _p, _i64, _u64, _i64, _p, _f, _i32, _i32, _p, _p, _p]),
    "dbx_bbb_step": (C.c_int, [_p, _p, _i64, _p, _p, _p, _p, _p, _p, C.c_uint64, _i64, C.c_float, C.c_float, C.c_int32, C.c_float, C.c_float,
                               C.c_float, C.c_float, C.c_int32, _p, _p, _p, _p]),
    "dbx_bbb_train_epoch": (C.c_int, [_p, _p, _p, _i64, _i64, _p, _p, _p, _p, C.c_uint64, _i64, C.c_float, C.c_float, C.c_int32, C.c_float, C.c_float,
                                      C.c_float, C.c_float, C.c_int32, _p, _p, _p, _p]),
    "dbx_bbb_step_ibp_mc": (C.c_int, [_p, _p, _i64, _p, _p, _p, _p, _p, _p, C.c_uint64, _i64, C.c_float, C.c_float, C.c_int32, _p, _p, _p, _p, _p]),
    "dbx_input_gradient": (C.c_int, [_p, _p, _i64, _p, _i64, _i64, C.c_uint64, _i64, _p, C.c_float, C.c_int32, _i64, _p, C.c_int32, _p, _p]),
    "dbx_input_gradient_one": (C.c_int, [_p, _p, _i64, _p, _p, _p, _p]),
    "dbx_count_predicate_fgsm": (C.c_int, [_p, _p, _i64, _p, _p, C.c_float, C.c_float, C.c_float, _i64, C.c_uint64, _i64, _p, C.c_float, C.c_int32, C.c_int32, _p, _p, _p]),
    "dbx_ibp_predict": (C.c_int, [_p, _p, _i64, C.c_float, C.c_int32, _p, _i64, C.c_uint64, _i64, _p, C.c_float, C.c_int32, _i64, C.c_int32,
                                  C.c_int32, _p, _p, _p, _p, _p, _p]),
    "dbx_predict_samples": (C.c_int, [_p, _p, _i64, _i64, _u64, _i64, _p, _f, _i32, _i32, _p, _p]),
    "dbx_fgsm_samples": (C.c_int, [_p, _p, _i64, _p, _f, _f, _f, _i64, _u64, _i64, _p, _f, C.c_int32, _p, _p]),
    "dbx_ibp_samples": (C.c_int, [_p, _p, _p, _i64, _f, _i32, _i64, _u64, _i64, _p, _f, _i32, _p, _p, _p]),
    "dbx_ibp_state": (C.c_int, [_p, _p, _p, _i64, _p, C.c_float, _p, _p, _p]),
    "dbx_ibp_one": (C.c_int, [_p, _p, _i64, C.c_float, C.c_int32, _p, C.c_int32, _p, _p, _p]),
    "dbx_upload_f32": (C.c_int, [_p, _p, _p, _i64, _p]),
    "dbx_comm_create": (C.c_int, [_i32, _i32, _i32, _i64, C.POINTER(_p)]),
    "dbx_comm_handle": (C.c_int, [_p, _p]),
    "dbx_comm_connect": (C.c_int, [_p, _p]),
    "dbx_comm_allreduce": (C.c_int, [_p, _p, _i64, _p, _i64, _p]),
    "dbx_comm_status": (C.c_int, [_p, C.POINTER(_i32), C.POINTER(_i64)]),
    "dbx_comm_destroy": (C.c_int, [_p]),
    "dbx_predict_host": (C.c_int, [_p, _p, _i64, _i64, _u64, _i64, _f, _i32, _i32, _p, _p]),
    "dbx_activation": (C.c_int, [_i32, _p, _i64, _i64, _p]),
    "dbx_philox_raw": (C.c_int, [_u64, _i64, _i64, _p, _p]),
}


def declared_symbols() -> list:
    """Every function the public header declares (used by the CPU-side ABI test)."""
    with open(HEADER_PATH) as f:
        src = f.read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dbx_[a-z0-9_]+)\s*\(", src)))


def load():
    """Load the shared library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DbxError(
            "libdeepbayes_b200.so not found at %s: build it with `python -m deepbayes_b200.build` "
            "(needs nvcc 12.9, targets sm_100a). There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def check(rc: int):
    if rc != 0:
        raise DbxError(load().dbx_last_error().decode("utf-8", "replace"))


def set_option(name: str, value: int) -> None:
    """Process-wide dispatch switch (`name` without the DBX_ prefix); see include/deepbayes_b200.h."""
    check(load().dbx_set_option(name.encode(), int(value)))


def get_option(name: str) -> int:
    v = _i64()
    check(load().dbx_get_option(name.encode(), C.byref(v)))
    return int(v.value)


def reset_options() -> None:
    check(load().dbx_reset_options())


def launch_count() -> int:
    return int(load().dbx_launch_count())
