# same export list as deepbayes/optimizers/__init__.py:2-10
from .optimizer import Optimizer  # noqa: F401
from .bayesbybackprop import BayesByBackprop  # noqa: F401
from .others import (NoisyAdam, VariationalOnlineGuassNewton, StochasticGradientDescent,  # noqa: F401
                     StochasticWeightAveragingGaussian, Adam, HamiltonianMonteCarlo, SGHamiltonianMonteCarlo)
SGHMC = SGHamiltonianMonteCarlo
