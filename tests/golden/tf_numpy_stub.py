"""NumPy stand-ins for the TensorFlow / Keras / TFP / statsmodels symbols that the reference's
hot-path source files touch, so that /root/reference/deepbayes/{posteriormodel.py,
optimizers/optimizer.py, optimizers/bayesbybackprop.py} can be EXECUTED UNMODIFIED in a
container without TensorFlow (see make_reference_fixtures.py).

Only leaf numerics are provided here (fp32 Dense forward, softmax, relu, softplus with TF's
thresholds, elementwise math, a seeded normal generator); every line of control flow that
runs -- RNG consumption order, var-used-as-std, fp32 running sum, /float(n), rho = log(exp(s)-1),
W = mean + softplus(rho)*noise -- is the reference's own.
"""
import json
import os
import sys
import types

import numpy as np

F32 = np.float32
_SP_THR = F32(np.log(np.finfo(np.float32).eps) + 2.0)


def _softplus(x):
    x = np.asarray(x, F32)
    with np.errstate(over="ignore"):
        ex = np.exp(x)
        mid = np.log1p(ex)
    return np.where(x > -_SP_THR, x, np.where(x < _SP_THR, ex, mid)).astype(F32)


def _act(name):
    def relu(z): return np.maximum(z, F32(0))
    def linear(z): return z
    def softmax(z):
        z = np.asarray(z, F32)
        m = z.max(axis=-1, keepdims=True)
        e = np.exp(z - m)
        return (e / e.sum(axis=-1, keepdims=True)).astype(F32)
    def tanh(z): return np.tanh(z)
    def sigmoid(z): return (1 / (1 + np.exp(-z))).astype(F32)
    f = {"relu": relu, "linear": linear, None: linear, "softmax": softmax, "tanh": tanh, "sigmoid": sigmoid}[name]
    f.__name__ = name or "linear"
    return f


class Dense:
    def __init__(self, units, activation=None, input_shape=None, **kw):
        self.units, self.activation, self.input_shape_arg = units, _act(activation), input_shape
        self.W = self.b = None

    def build(self, fan_in, rng):
        lim = np.sqrt(6.0 / (fan_in + self.units))
        self.W = rng.uniform(-lim, lim, (fan_in, self.units)).astype(F32)
        self.b = np.zeros(self.units, F32)

    def get_weights(self): return [self.W, self.b]

    def __call__(self, x):
        return self.activation(np.asarray(x, F32) @ self.W + self.b)


class _Weight:
    def __init__(self, a): self.a = a; self.shape = a.shape


class Sequential:
    def __init__(self, layers=None):
        self.layers = []
        for l in (layers or []): self.add(l)

    def add(self, l):
        fan_in = l.input_shape_arg[-1] if l.input_shape_arg is not None else self.layers[-1].units
        l.build(fan_in, np.random.RandomState(len(self.layers)))
        self.layers.append(l)

    @property
    def weights(self): return [_Weight(w) for w in self.get_weights()]

    @property
    def trainable_variables(self): return self.get_weights()

    def get_weights(self): return [w for l in self.layers for w in l.get_weights()]

    def set_weights(self, ws):
        ws = [np.asarray(w, dtype=F32) for w in ws]   # Keras casts to the variable dtype (fp32)
        for i, l in enumerate(self.layers):
            l.W, l.b = ws[2 * i], ws[2 * i + 1]

    def __call__(self, x):
        h = np.asarray(x, F32)
        for l in self.layers: h = l(h)
        return h

    def summary(self): return None
    def save(self, path): pass
    def to_json(self): return json.dumps({"stub": True})


def _load_model(path):
    d = os.path.dirname(path)
    cfg = json.load(open(os.path.join(d, "arch.json")))["config"]["layers"]
    layers, shape = [], None
    for lc in cfg:
        c = lc["config"]
        if c.get("batch_input_shape") is not None and shape is None: shape = c["batch_input_shape"][1:]
        if lc["class_name"] == "Dense":
            layers.append(Dense(c["units"], c["activation"], input_shape=shape if not layers else None))
    return Sequential(layers)


class _Tape:
    def __init__(self, persistent=False): pass
    def __enter__(self): return self
    def __exit__(self, *a): return False
    def watch(self, v): pass
    def gradient(self, loss, vs): return [np.zeros_like(np.asarray(v, F32)) for v in vs]


class _Metric:
    def __init__(self, name=None): self.name, self.calls = name, []
    def __call__(self, *a): self.calls.append(a)
    def result(self): return 0.0
    def reset_states(self): pass


def install(noise_log):
    """Put the stand-in modules into sys.modules.  noise_log collects every tf.random.normal draw."""
    tf = types.ModuleType("tensorflow")
    tf.float32 = F32
    tf.zeros = lambda shape, dtype=None: np.zeros(tuple(shape), F32)
    tf.ones = lambda shape, dtype=None: np.ones(tuple(shape), F32)
    tf.multiply = lambda a, b: (np.asarray(a, F32) * np.asarray(b, F32)).astype(F32)
    tf.matmul = lambda a, b: np.asarray(a, F32) @ np.asarray(b, F32)
    tf.cast = lambda a, dtype=None: np.asarray(a, F32)
    tf.convert_to_tensor = lambda a, dtype=None: np.asarray(a, F32)
    tf.squeeze = np.squeeze
    tf.GradientTape = _Tape
    m = types.ModuleType("tensorflow.math")
    m.softplus = _softplus
    m.log = lambda a: np.log(np.asarray(a, F32)).astype(F32)
    m.exp = lambda a: np.exp(np.asarray(a, F32)).astype(F32)
    m.add = lambda a, b: (np.asarray(a, F32) + np.asarray(b, F32)).astype(F32)
    m.subtract = lambda a, b: (np.asarray(a, F32) - np.asarray(b, F32)).astype(F32)
    m.multiply = tf.multiply
    m.divide = lambda a, b: (np.asarray(a, F32) / np.asarray(b, F32)).astype(F32)
    tf.math = m
    r = types.ModuleType("tensorflow.random")

    def normal(shape, mean=0.0, stddev=1.0, **kw):
        z = np.random.standard_normal(tuple(shape)).astype(F32)   # seeded by the caller
        noise_log.append(z)
        return (np.asarray(mean, F32) + F32(stddev) * z).astype(F32)
    r.normal = normal
    tf.random = r
    keras = types.ModuleType("tensorflow.keras")
    models = types.ModuleType("tensorflow.keras.models")
    models.load_model = _load_model
    models.Sequential = Sequential
    layers = types.ModuleType("tensorflow.keras.layers")
    layers.Dense = Dense
    metrics = types.ModuleType("tensorflow.keras.metrics")
    metrics.Mean = _Metric
    metrics.SparseCategoricalAccuracy = _Metric
    metrics.RootMeanSquaredError = _Metric
    losses = types.ModuleType("tensorflow.keras.losses")
    losses.SparseCategoricalCrossentropy = lambda *a, **k: (lambda y, p: 0.0)
    losses.mean_squared_error = lambda y, p: 0.0
    losses.MeanAbsoluteError = lambda *a, **k: (lambda y, p: 0.0)
    losses.MeanSquaredError = lambda *a, **k: (lambda y, p: 0.0)
    losses.__getattr__ = lambda name: (lambda *a, **k: (lambda y, p: 0.0))
    losses.sparse_categorical_crossentropy = lambda y, p: 0.0
    keras.models, keras.layers, keras.metrics, keras.losses = models, layers, metrics, losses
    tf.keras = keras
    # the pickled EagerTensors in mean.npy / var.npy name this symbol
    tpy = types.ModuleType("tensorflow.python"); tfw = types.ModuleType("tensorflow.python.framework")
    ops = types.ModuleType("tensorflow.python.framework.ops")
    ops.convert_to_tensor = lambda x, *a, **k: np.asarray(x)
    tfp = types.ModuleType("tensorflow_probability")
    dist = types.ModuleType("tensorflow_probability.distributions")
    dist.Exponential = lambda rate=1.0: types.SimpleNamespace(sample=lambda: rate)
    dist.Normal = lambda *a, **k: types.SimpleNamespace(log_prob=lambda x: np.zeros_like(np.asarray(x, F32)))
    tfp.distributions = dist
    sm = types.ModuleType("statsmodels"); sms = types.ModuleType("statsmodels.stats")
    smp = types.ModuleType("statsmodels.stats.proportion")
    smp.proportion_confint = lambda *a, **k: (0.0, 1.0)
    for name, mod in {"tensorflow": tf, "tensorflow.math": m, "tensorflow.random": r, "tensorflow.keras": keras,
                      "tensorflow.keras.models": models, "tensorflow.keras.layers": layers,
                      "tensorflow.keras.metrics": metrics, "tensorflow.keras.losses": losses,
                      "tensorflow.python": tpy, "tensorflow.python.framework": tfw,
                      "tensorflow.python.framework.ops": ops, "tensorflow_probability": tfp,
                      "tensorflow_probability.distributions": dist, "statsmodels": sm, "statsmodels.stats": sms,
                      "statsmodels.stats.proportion": smp}.items():
        sys.modules[name] = mod
    return tf
