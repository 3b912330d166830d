"""Optimizer base -- the inference-side surface of deepbayes/optimizers/optimizer.py:
``compile`` (:33-97), priors (:204-257), ``sample`` (:197-202), ``predict`` (:293-298),
``save`` (:259-276), ``load`` (:279-290).  The training loop (train / model_validate /
logging, :100-194) is outside the hot path named by BASELINE.json and is not built; ``train``
raises.  All tensors are NumPy on the host and one flat fp32 vector on the device.
"""
from __future__ import annotations

import math
from abc import ABC, abstractmethod

import numpy as np

from .. import formats
from ..engine import Engine, flatten_weights
from ..keras_compat import from_keras


def softplus(x):
    """tf.math.softplus with TensorFlow's thresholds (optimizer.py:25-26).
    Host copy used only for values written to disk / returned to the caller."""
    x = np.asarray(x, dtype=np.float32)
    thr = np.float32(np.log(np.finfo(np.float32).eps) + 2.0)
    with np.errstate(over="ignore"):
        ex = np.exp(x)
        mid = np.log1p(ex)
    return np.where(x > -thr, x, np.where(x < thr, ex, mid)).astype(np.float32)


class Optimizer(ABC):
    def __init__(self):
        # the reference prints a reminder here (optimizer.py:29-30); kept quiet on purpose
        self._engine = None
        self._rho_param = False

    # ------------------------------------------------------------------ compile
    @abstractmethod
    def compile(self, keras_model, loss_fn, batch_size, learning_rate, decay, epochs, prior_mean, prior_var, **kwargs):
        self.mode = kwargs.get('mode', 'classification')
        self.classes = kwargs.get('classes', 10)
        self.model = from_keras(keras_model)
        self.batch_size = batch_size
        self.learning_rate = learning_rate
        self.decay = decay
        self.epochs = epochs
        self.loss_func = loss_fn
        self.log_dir = kwargs.get('log_file', '/tmp/BayesKeras.log')
        self.det = kwargs.get('deterministic', False)
        self.inflate_prior = kwargs.get('inflate_prior', 1)
        self.input_noise = kwargs.get('input_noise', 0.0)
        # prior and posterior are generated separately: the posterior mean starts at the
        # PRIOR mean, not at the Keras initialiser (optimizer.py:51-52)
        self.prior_mean, self.prior_var = self.prior_generator(prior_mean, prior_var)
        self.posterior_mean, self.posterior_var = self.prior_generator(prior_mean, prior_var)
        self.robust_train = kwargs.get('robust_train', 0)
        self.epsilon = kwargs.get('epsilon', 0.1)
        self.robust_lambda = kwargs.get('rob_lam', 0.5)
        self.robust_linear = kwargs.get('linear_schedule', True)
        if self.robust_linear:
            self.epochs += 1  # optimizer.py:77-79 (Appendix A.7)
        self.attack_loss = kwargs.get('attack_loss', None)
        self.loss_monte_carlo = kwargs.get('loss_mc', 2)
        self.input_upper = math.inf
        self.input_lower = -math.inf
        self.acc_log, self.rob_log, self.loss_log = [], [], []
        # device-side extras (not in the reference)
        self.compute_mode = kwargs.get('compute_mode', 'fp32')
        self.train_mode = kwargs.get('train_mode', 'fp32')    # arithmetic of the training step's GEMMs (not in the reference)
        self.seed = int(kwargs.get('seed', np.random.randint(0, 2 ** 31 - 1)))
        self._next_sample = 0
        self._device = kwargs.get('device', None)
        self._engine = None
        return self

    # ------------------------------------------------------------------ priors
    def _gen_implicit_prior(self):
        """optimizer.py:204-228."""
        prior_mean, prior_var = [], []
        for layer in self.model.layers:
            ws = layer.get_weights()
            if len(ws) < 2:
                continue
            sha, b_sha = ws[0].shape, ws[1].shape
            if len(sha) > 2:
                nin = 1
                for i in range(len(sha) - 1):
                    nin *= sha[i]
            else:
                nin = sha[0]
            std = math.sqrt(self.inflate_prior / nin)
            prior_mean += [np.zeros(sha, np.float32), np.zeros(b_sha, np.float32)]
            prior_var += [np.ones(sha, np.float32) * np.float32(std), np.ones(b_sha, np.float32) * np.float32(std)]
        return prior_mean, prior_var

    def prior_generator(self, means, vars):
        """optimizer.py:230-257."""
        if type(means) == int and type(vars) == int:
            if means < 0 or vars < 0:
                return self._gen_implicit_prior()
        shapes = self.model.weight_shapes()
        if type(means) == int or type(means) == float:
            if means == -1:
                means = 0.0
            means = [means for _ in range(len(shapes))]
        if type(vars) == int or type(vars) == float:
            if vars == -1:
                vars = 0.0
            vars = [vars for _ in range(len(shapes))]
        model_mean, model_var = [], []
        index = 0.0
        for shp in shapes:
            param_index = math.floor(index / 2.0)
            model_mean.append(np.ones(shp, np.float32) * np.float32(means[param_index]))
            model_var.append(np.ones(shp, np.float32) * np.float32(vars[param_index]))
            index += 1
        return model_mean, model_var

    # ------------------------------------------------------------------ device state
    def _sigma_for_device(self):
        """(array, is_rho) handed to the device as the sampling scale."""
        return self.posterior_var, self._rho_param

    def engine(self) -> Engine:
        if self._engine is None:
            self._engine = Engine(self.model, self._device)
            self._synced = None
        key = (id(self.posterior_mean), id(self.posterior_var))
        if self._synced != key:
            sig, is_rho = self._sigma_for_device()
            self._engine.set_gaussian(self.posterior_mean, sig, rho=is_rho)
            self._synced = key
        return self._engine

    def _take_ids(self, n):
        s0 = self._next_sample
        self._next_sample += n
        return s0

    # ------------------------------------------------------------------ sample / predict
    def sample(self):
        """optimizer.py:197-202: N(posterior_mean, scale=posterior_var) -- var used as std."""
        eng = self.engine()
        W = eng.sample(1, self.seed, self._take_ids(1))[0].cpu().numpy()
        out, o = [], 0
        for s in self.model.weight_shapes():
            k = int(np.prod(s))
            out.append(W[o:o + k].reshape(s).astype(np.float64))     # np.random.normal returns float64 (optimizer.py:199-201)
            o += k
        return out

    def predict(self, input, n=1):
        """optimizer.py:293-298: ONE forward with the current model weights, no sampling,
        ``n`` ignored (Appendix A.4)."""
        return self.model(input)

    def _predict(self, input):
        return self.model(input)

    def set_weights(self, weights):
        self.model.set_weights(weights)

    @abstractmethod
    def train(self, X_train, y_train, X_test=None, y_test=None, **kwargs):
        raise NotImplementedError(
            "deepbayes_b200 builds the posterior-predictive / reparameterised-forward path only; "
            "the training loop (optimizer.py:100-133) is out of scope (SURVEY.md section 8(f))")

    # ------------------------------------------------------------------ save / load
    def save(self, path):
        """optimizer.py:259-276: mean.npy, var.npy (raw posterior_var), info.pkl with every
        int/float attribute, arch.json (+ weights npz in place of model.h5)."""
        self.info = {k: v for k, v in self.__dict__.items() if type(v) == int or type(v) == float}
        formats.save_gaussian_posterior(path, self.model, self.posterior_mean, self.posterior_var, self.info)

    def load(self, path):
        """optimizer.py:279-290: reads mean/var and stores rho = log(exp(var)-1) into
        ``posti_var`` / ``posti_mean`` (the attribute names the reference uses)."""
        posti_mean = formats.load_tensor_list(path + "/mean.npy")
        posti_var = formats.load_tensor_list(path + "/var.npy")
        self.posti_var = [np.log(np.exp(v) - np.float32(1)).astype(np.float32) for v in posti_var]
        self.posti_mean = posti_mean
