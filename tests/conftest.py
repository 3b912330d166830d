import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with `pytest -m gpu` under gpurun)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as o
    return o


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture
def dbx_opt():
    """Setter for the library's process-wide dispatch switches (include/deepbayes_b200.h: dbx_set_option); every switch is
    back at its default after the test.  ``dbx_opt("NO_FUSED", 1)``; value None = that option's default."""
    from deepbayes_b200 import _lib
    touched = {}

    def set_(name, value):
        if value is None:
            touched.pop(name, None)
            _lib.reset_options()
            for k, v in touched.items():
                _lib.set_option(k, v)
        else:
            touched[name] = int(value)
            _lib.set_option(name, int(value))

    yield set_
    _lib.reset_options()
