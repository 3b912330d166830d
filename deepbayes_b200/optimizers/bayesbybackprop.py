"""BayesByBackprop -- hot-path part of deepbayes/optimizers/bayesbybackprop.py:
``compile`` (:31-44, rho = log(exp(std)-1)), the sampling + forward half of ``step``
(:46-65, W = mu + softplus(rho)*eps then model(features)), ``sample`` (:176-181) and ``save``
(:184-195).  The backward / SGD update (:138-170) is the next row.
"""
from __future__ import annotations

import numpy as np

from .. import formats
from . import optimizer
from .optimizer import softplus


class BayesByBackprop(optimizer.Optimizer):
    def __init__(self):
        super().__init__()

    def compile(self, keras_model, loss_fn=None, batch_size=64, learning_rate=0.15, decay=0.0,
                epochs=10, prior_mean=-1, prior_var=-1, **kwargs):
        super().compile(keras_model, loss_fn, batch_size, learning_rate, decay,
                        epochs, prior_mean, prior_var, **kwargs)
        # posterior_var now holds rho (bayesbybackprop.py:39-40)
        with np.errstate(divide="ignore"):
            self.posterior_var = [np.log(np.exp(v) - np.float32(1)).astype(np.float32) for v in self.posterior_var]
        self._rho_param = True
        self.kl_weight = kwargs.get('kl_weight', 1.0)
        return self

    def forward_sample(self, features, eps=None, return_weights=False):
        """bayesbybackprop.py:50-65 on the device: noise ~ N(0,1) per weight (or ``eps``
        injected, a list [eps_W0, eps_b0, ...] / flat vector), W = mean + softplus(rho)*noise
        formed in fp32 exactly in that order, forward of the mini-batch.  Returns predictions
        (and, if asked, the sampled weights and ``noise_used``, :57)."""
        eng = self.engine()
        _, oshape = self.model.rows_and_outshape(np.asarray(features))
        e = None if eps is None else eng._as_flat(eps)
        step = self._take_ids(1)
        if return_weights:
            out, W, noise = eng.bbb_forward(features, self.seed, step, e, self.compute_mode, True)
            Wn, nn = W.cpu().numpy(), noise.cpu().numpy()
            ws, ns, o = [], [], 0
            for s in self.model.weight_shapes():
                k = int(np.prod(s))
                ws.append(Wn[o:o + k].reshape(s)); ns.append(nn[o:o + k].reshape(s)); o += k
            self.model.set_weights(ws)  # :59
            return out.cpu().numpy().reshape(oshape), ws, ns
        out = eng.bbb_forward(features, self.seed, step, e, self.compute_mode, False)
        return out.cpu().numpy().reshape(oshape)

    def step(self, features, labels, lrate):
        raise NotImplementedError(
            "BayesByBackprop.step: only the sampling + forward half (forward_sample) is built; the "
            "reparameterised backward and SGD update (bayesbybackprop.py:138-170) are SURVEY.md 8(f) rank 2")

    def train(self, X_train, y_train, X_test=None, y_test=None, **kwargs):
        return super().train(X_train, y_train, X_test, y_test, **kwargs)

    def sample(self):
        """bayesbybackprop.py:176-181: N(mean, scale=softplus(rho)) -- the device holds
        sigma = softplus(rho) (dbx_set_gaussian_rho)."""
        return super().sample()

    def save(self, path):
        """bayesbybackprop.py:184-195: var.npy = softplus(rho) (a std); no info.pkl."""
        var = [softplus(v) for v in self.posterior_var]
        formats.save_gaussian_posterior(path, self.model, self.posterior_mean, var, None)
