"""The other exported optimizer names (deepbayes/optimizers/__init__.py:2-10).  Their training
steps are outside the hot path; what the posterior-predictive
path needs from them -- ``compile`` (posterior init), ``sample``, ``predict`` and the ``save``
format that PosteriorModel consumes -- is provided so a saved posterior of every family
can be produced and reloaded.
"""
from __future__ import annotations

import numpy as np

from .. import formats
from . import optimizer


class _InferenceOnly(optimizer.Optimizer):
    def __init__(self):
        super().__init__()

    def compile(self, keras_model, loss_fn=None, batch_size=64, learning_rate=0.15, decay=0.0,
                epochs=10, prior_mean=-1, prior_var=-1, **kwargs):
        super().compile(keras_model, loss_fn, batch_size, learning_rate, decay, epochs, prior_mean, prior_var, **kwargs)
        return self

    def train(self, X_train, y_train, X_test=None, y_test=None, **kwargs):
        return super().train(X_train, y_train, X_test, y_test, **kwargs)


class VariationalOnlineGuassNewton(_InferenceOnly):
    """blrvi.py: compile adds beta_1/beta_2/lam/N (:31-45); save writes var = 1/(N*s)
    (:171-180), which PosteriorModel then uses as a std (Appendix A.1)."""

    def compile(self, keras_model, loss_fn=None, batch_size=64, learning_rate=0.15, decay=0.0,
                epochs=10, prior_mean=-1, prior_var=-1, **kwargs):
        super().compile(keras_model, loss_fn, batch_size, learning_rate, decay, epochs, prior_mean, prior_var, **kwargs)
        self.beta_1 = kwargs.get('beta_1', 0.999)
        self.beta_2 = kwargs.get('beta_2', 0.9999)
        self.lam = kwargs.get('lam', 1.0)
        self.N = kwargs.get('N', 60000)
        return self


class NoisyAdam(VariationalOnlineGuassNewton):
    pass


class StochasticGradientDescent(_InferenceOnly):
    pass


class Adam(_InferenceOnly):
    def sample(self):
        """adam.py:159-160: returns the current weights."""
        return self.model.get_weights()


class StochasticWeightAveragingGaussian(_InferenceOnly):
    """swag.py: posterior = per-weight mean / std of the recorded iterates (:164-176)."""

    def compile(self, *a, **kw):
        super().compile(*a, **kw)
        self.weights_stack = []
        return self

    def record(self, weights):
        self.weights_stack.append([np.asarray(w, np.float32) for w in weights])

    def get_posterior(self):
        n_t = len(self.weights_stack[0])
        mean = [np.mean(np.stack([w[i] for w in self.weights_stack]), axis=0) for i in range(n_t)]
        var = [np.std(np.stack([w[i] for w in self.weights_stack]), axis=0) for i in range(n_t)]
        self.posterior_mean, self.posterior_var = mean, var

    def save(self, path):
        self.get_posterior()
        formats.save_gaussian_posterior(path, self.model, self.posterior_mean, self.posterior_var, None)


class HamiltonianMonteCarlo(_InferenceOnly):
    """hmc.py: the chain itself is training; ``posterior_samples`` / ``num_rets`` and the
    ``save`` layout (:308-322) are what the stored-posterior path consumes."""

    def compile(self, *a, **kw):
        super().compile(*a, **kw)
        self.posterior_samples = []
        self.num_rets = [0]
        return self

    def record(self, weights, multiplicity=1):
        self.posterior_samples.append([np.asarray(w, np.float32) for w in weights])
        self.num_rets.append(multiplicity)

    def sample(self):
        """posteriormodel.py:77 semantics on the live chain: a stored state ~ multiplicity."""
        f = np.asarray(self.num_rets[1:] if self.num_rets[0] == 0 else self.num_rets, dtype=np.float64)
        i = np.random.choice(range(len(f)), p=f / f.sum())
        return [w.copy() for w in self.posterior_samples[i]]

    def save(self, path):
        if self.num_rets[0] == 0:
            self.num_rets = self.num_rets[1:]
        formats.save_sample_posterior(path, self.model, self.posterior_mean, self.posterior_samples, self.num_rets)


class SGHamiltonianMonteCarlo(HamiltonianMonteCarlo):
    pass
