"""deepbayes_b200 -- B200-native (sm_100a) implementation of the deepbayes posterior-predictive
hot path behind the reference's Python surface:

    from deepbayes_b200 import PosteriorModel, optimizers, analyzers
    opt = optimizers.BayesByBackprop().compile(model, loss_fn=None, ...)
    PosteriorModel(path).predict(x, n=1024)

Importing the package never touches CUDA; any compute call needs libdeepbayes_b200.so (built
by ``python -m deepbayes_b200.build``) and a B200 -- there is no CPU fallback.
"""
from .keras_compat import Sequential, Dense, Conv2D, MaxPooling2D, Flatten, InputLayer, model_from_json  # noqa: F401
from .posteriormodel import PosteriorModel  # noqa: F401  (deepbayes/__init__.py:4)
from . import optimizers  # noqa: F401
from . import analyzers  # noqa: F401

__version__ = "0.1.0"
