# the reference star-imports attacks, verifiers, uncertainty, probverification
# (deepbayes/analyzers/__init__.py:1-4); only the hot-path callers are built here.
from .verifiers import *  # noqa: F401,F403
from .uncertainty import *  # noqa: F401,F403
from .attacks import gradient_expectation, zeroth_order_gradient, FGSM, PGD, _PGD, CW  # noqa: F401
from . import verifiers, uncertainty, attacks  # noqa: F401
