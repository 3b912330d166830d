"""Minimal Keras-shaped model descriptors.

The reference hands a ``tf.keras`` Sequential to ``Optimizer.compile`` (optimizer.py:33-40)
and ``PosteriorModel`` reloads one (posteriormodel.py:37,50); the analyzers then use
``model.layers[i]``, ``.activation``, ``.get_weights()``, ``.set_weights()``, ``model(x)``
(verifiers.py:75-94, 435-438).  TensorFlow is not available on the B200 image, so these
classes carry exactly that surface; their ``__call__`` runs on the GPU through the C ABI.
A real Keras model (when TF is importable) is accepted anywhere via ``from_keras``.
"""
from __future__ import annotations

import json
import math
from typing import List, Optional, Sequence

import numpy as np

from . import _lib


class Activation:
    """Callable activation object (``layer.activation`` in Keras)."""

    def __init__(self, name):
        name = "linear" if name is None else (name if isinstance(name, str) else getattr(name, "__name__", str(name)))
        if name not in _lib.ACT:
            raise ValueError("unsupported activation %r" % (name,))
        self.__name__ = self.name = name

    def __call__(self, x):
        from .engine import device_activation
        return device_activation(self.name, x)

    def __repr__(self):
        return "Activation(%s)" % self.name


class Layer:
    kind = None
    has_weights = False

    def __init__(self, name=None):
        self.name = name or self.__class__.__name__.lower()
        self._weights: List[np.ndarray] = []
        self._model = None
        self.activation = Activation("linear")
        self.input_shape_arg = None

    def get_weights(self):
        if self._model is not None:
            self._model._materialise()
        return [w.copy() for w in self._weights]

    def set_weights(self, ws):
        ws = [np.asarray(w, dtype=np.float32) for w in ws]
        if len(ws) != len(self._weights) or any(a.shape != b.shape for a, b in zip(ws, self._weights)):
            raise ValueError("layer %s: weight shapes %s do not match %s" % (
                self.name, [w.shape for w in ws], [w.shape for w in self._weights]))
        self._weights = ws
        if self._model is not None:
            self._model._lazy = None

    @property
    def weights(self):
        return self.get_weights()

    def __call__(self, x):
        """Forward of this single layer on the GPU (posteriormodel.py:106-108,124-127)."""
        if self._model is None:
            raise RuntimeError("layer is not part of a built model")
        return self._model._call_layer(self, x)


class InputLayer(Layer):
    kind = "input"

    def __init__(self, input_shape=None, batch_input_shape=None, name=None):
        super().__init__(name)
        if batch_input_shape is not None:
            input_shape = tuple(batch_input_shape[1:])
        self.input_shape_arg = tuple(input_shape)


class Dense(Layer):
    kind = "dense"
    has_weights = True

    def __init__(self, units, activation=None, input_shape=None, use_bias=True, name=None, **_):
        super().__init__(name)
        if not use_bias:
            raise ValueError("use_bias=False is not supported (the reference assumes W,b pairs, optimizer.py:208-209)")
        self.units = int(units)
        self.activation = Activation(activation)
        self.input_shape_arg = tuple(input_shape) if input_shape is not None else None


class Conv2D(Layer):
    kind = "conv2d"
    has_weights = True

    def __init__(self, filters, kernel_size, strides=(1, 1), padding="valid", activation=None, input_shape=None,
                 use_bias=True, name=None, **_):
        super().__init__(name)
        if not use_bias:
            raise ValueError("use_bias=False is not supported")
        ks = (kernel_size, kernel_size) if isinstance(kernel_size, int) else tuple(kernel_size)
        st = (strides, strides) if isinstance(strides, int) else tuple(strides)
        self.filters, self.kernel_size, self.strides = int(filters), (int(ks[0]), int(ks[1])), (int(st[0]), int(st[1]))
        self.padding = padding.lower()
        self.activation = Activation(activation)
        self.input_shape_arg = tuple(input_shape) if input_shape is not None else None


class MaxPooling2D(Layer):
    kind = "maxpool2d"

    def __init__(self, pool_size=(2, 2), strides=None, padding="valid", name=None, **_):
        super().__init__(name)
        ps = (pool_size, pool_size) if isinstance(pool_size, int) else tuple(pool_size)
        st = ps if strides is None else ((strides, strides) if isinstance(strides, int) else tuple(strides))
        if padding.lower() != "valid":
            raise ValueError("MaxPooling2D: only padding='valid'")
        self.pool_size, self.strides = (int(ps[0]), int(ps[1])), (int(st[0]), int(st[1]))


class Flatten(Layer):
    kind = "flatten"


class Sequential:
    """Keras-Sequential-shaped container.  ``layers`` excludes the InputLayer, as in Keras."""

    def __init__(self, layers: Optional[Sequence[Layer]] = None, name="sequential"):
        self.name = name
        self.layers: List[Layer] = []
        self.input_shape: Optional[tuple] = None  # per-sample shape, e.g. (784,), (1,784), (32,32,3)
        self.built = False
        self._engine = None
        self._lazy = None  # (callable) pending weight materialisation after predict()
        self._sub_engines = {}
        for l in (layers or []):
            self.add(l)

    # ---- construction -----------------------------------------------------------------
    def add(self, layer: Layer):
        if isinstance(layer, InputLayer):
            self.input_shape = layer.input_shape_arg
            return
        if self.input_shape is None and layer.input_shape_arg is not None:
            self.input_shape = layer.input_shape_arg
        layer._model = self
        self.layers.append(layer)
        self.built = False
        if self.input_shape is not None:
            self.build()

    def build(self, input_shape=None, seed=None):
        if input_shape is not None:
            self.input_shape = tuple(input_shape[1:]) if (len(input_shape) and input_shape[0] is None) else tuple(input_shape)
        if self.input_shape is None:
            raise ValueError("input shape unknown: give input_shape= to the first layer or call build()")
        rng = np.random.RandomState(seed)
        first = next((l for l in self.layers if l.has_weights), None)
        if first is None:
            raise ValueError("model has no weighted layer")
        shp = tuple(self.input_shape)
        if first.kind == "conv2d":
            if len(shp) != 3:
                raise ValueError("Conv2D-first models need an (H,W,C) input shape, got %r" % (shp,))
            cur = shp
        else:
            cur = (int(shp[-1]),)  # Dense acts on the last axis (Keras)
        self.feature_shape = cur
        for l in self.layers:
            if l.kind == "dense":
                if len(cur) != 1:
                    raise ValueError("Dense after a spatial layer needs a Flatten first")
                shapes = [(cur[0], l.units), (l.units,)]
                cur = (l.units,)
            elif l.kind == "conv2d":
                kh, kw = l.kernel_size
                shapes = [(kh, kw, cur[2], l.filters), (l.filters,)]
                if l.padding == "same":
                    ho, wo = -(-cur[0] // l.strides[0]), -(-cur[1] // l.strides[1])
                else:
                    ho, wo = (cur[0] - kh) // l.strides[0] + 1, (cur[1] - kw) // l.strides[1] + 1
                cur = (ho, wo, l.filters)
            elif l.kind == "maxpool2d":
                cur = ((cur[0] - l.pool_size[0]) // l.strides[0] + 1, (cur[1] - l.pool_size[1]) // l.strides[1] + 1, cur[2])
                shapes = []
            elif l.kind == "flatten":
                cur = (int(np.prod(cur)),)
                shapes = []
            else:
                raise ValueError("unsupported layer kind %r" % l.kind)
            if shapes and (len(l._weights) != 2 or l._weights[0].shape != shapes[0]):
                fan_in = int(np.prod(shapes[0][:-1]))
                fan_out = shapes[0][-1] * (int(np.prod(shapes[0][:-2])) if len(shapes[0]) > 2 else 1)
                lim = math.sqrt(6.0 / (fan_in + fan_out))  # GlorotUniform, the Keras default (arch.json)
                l._weights = [rng.uniform(-lim, lim, shapes[0]).astype(np.float32), np.zeros(shapes[1], np.float32)]
        self.output_dim = int(cur[0])
        self.built = True
        self._engine = None
        self._sub_engines = {}

    # ---- weights ----------------------------------------------------------------------
    def _materialise(self):
        if self._lazy is not None:
            fn, self._lazy = self._lazy, None
            self._assign(fn())

    def _assign(self, ws):
        i = 0
        for l in self.layers:
            if l.has_weights:
                l._weights = [np.asarray(ws[i], np.float32).reshape(l._weights[0].shape),
                              np.asarray(ws[i + 1], np.float32).reshape(l._weights[1].shape)]
                i += 2

    def get_weights(self):
        self._materialise()
        out = []
        for l in self.layers:
            out.extend(w.copy() for w in l._weights)
        return out

    def set_weights(self, ws):
        ws = list(ws)
        n = sum(2 for l in self.layers if l.has_weights)
        if len(ws) != n:
            raise ValueError("set_weights: expected %d arrays, got %d" % (n, len(ws)))
        self._lazy = None
        i = 0
        for l in self.layers:
            if l.has_weights:
                l.set_weights(ws[i:i + 2])
                i += 2

    def set_weights_lazy(self, fn):
        """predict() leaves the model holding the last sampled weights (posteriormodel.py:96);
        the device draws them, so the host copy is fetched only if somebody looks."""
        self._lazy = fn

    @property
    def weights(self):
        return self.get_weights()

    @property
    def trainable_variables(self):
        return self.get_weights()

    def weight_shapes(self):
        out = []
        for l in self.layers:
            out.extend(w.shape for w in l._weights)
        return out

    def count_params(self):
        return int(sum(int(np.prod(s)) for s in self.weight_shapes()))

    # ---- (de)serialisation ------------------------------------------------------------
    def to_json(self):
        cfg = [{"class_name": "InputLayer", "config": {"batch_input_shape": [None] + list(self.input_shape),
                                                       "dtype": "float32", "name": "input"}}]
        for l in self.layers:
            c = {"name": l.name, "dtype": "float32"}
            if l.kind == "dense":
                c.update(units=l.units, activation=l.activation.name, use_bias=True)
                cls = "Dense"
            elif l.kind == "conv2d":
                c.update(filters=l.filters, kernel_size=list(l.kernel_size), strides=list(l.strides),
                         padding=l.padding, activation=l.activation.name, use_bias=True, data_format="channels_last")
                cls = "Conv2D"
            elif l.kind == "maxpool2d":
                c.update(pool_size=list(l.pool_size), strides=list(l.strides), padding="valid", data_format="channels_last")
                cls = "MaxPooling2D"
            else:
                cls = "Flatten"
            cfg.append({"class_name": cls, "config": c})
        return json.dumps({"class_name": "Sequential", "config": {"name": self.name, "layers": cfg},
                           "keras_version": "deepbayes_b200", "backend": "cuda-sm_100a"})

    def save(self, path):
        """Stands in for ``model.save(path+'/model.h5')`` (optimizer.py:273): HDF5 cannot be
        written without h5py, so the weights go to ``<path minus .h5>.weights.npz``."""
        base = path[:-3] if path.endswith(".h5") else path
        np.savez(base + ".weights.npz", *self.get_weights())

    def summary(self):
        lines = ["Model: %s  input %s" % (self.name, self.input_shape)]
        for l in self.layers:
            lines.append("  %-14s %-10s params %d" % (l.name, l.kind, sum(int(w.size) for w in l._weights)))
        lines.append("Total params: %d" % self.count_params())
        s = "\n".join(lines)
        print(s)
        return None

    # ---- device forward ---------------------------------------------------------------
    def engine(self):
        if self._engine is None:
            from .engine import Engine
            self._engine = Engine(self)
        return self._engine

    def rows_and_outshape(self, x: np.ndarray):
        """Flatten leading dims into rows the way Keras Dense does."""
        x = np.asarray(x)
        nfeat = len(self.feature_shape)
        if x.ndim < nfeat + 1 and not (nfeat == 1 and x.ndim >= 1):
            raise ValueError("input of shape %r does not match model input %r" % (x.shape, self.input_shape))
        lead = x.shape[:x.ndim - nfeat]
        if tuple(x.shape[x.ndim - nfeat:]) != tuple(self.feature_shape):
            raise ValueError("input of shape %r does not match model feature shape %r" % (x.shape, self.feature_shape))
        rows = int(np.prod(lead)) if len(lead) else 1
        has_flatten = any(l.kind == "flatten" for l in self.layers)
        out_lead = lead[:1] if (has_flatten and len(lead) > 1) else lead
        if has_flatten and len(lead) > 1 and nfeat == 1:
            raise ValueError("Flatten over extra leading dims is not supported for Dense-first models; "
                             "reshape the input to (batch, features)")
        return rows, tuple(out_lead) + (self.output_dim,)

    def __call__(self, x):
        """Single forward with the current weights (``self.model(input)``, posteriormodel.py:93)."""
        return self.engine().forward(x, self.get_weights())

    def predict(self, x):
        return self(x)

    def _call_layer(self, layer, x):
        idx = self.layers.index(layer)
        key = idx
        if key not in self._sub_engines:
            x_shape = tuple(np.asarray(x).shape[1:])
            sub = Sequential(name="sub%d" % idx)
            # a one-layer model sharing this layer's hyper-parameters
            clone = _clone_layer(layer)
            sub.input_shape = self._shape_before(idx)
            sub.add(clone)
            sub.build()
            self._sub_engines[key] = sub
        sub = self._sub_engines[key]
        if layer.has_weights:
            sub.set_weights(layer.get_weights())
        return sub(x)

    def _shape_before(self, idx):
        cur = tuple(self.feature_shape)
        for l in self.layers[:idx]:
            if l.kind == "dense":
                cur = (l.units,)
            elif l.kind == "conv2d":
                kh, kw = l.kernel_size
                if l.padding == "same":
                    cur = (-(-cur[0] // l.strides[0]), -(-cur[1] // l.strides[1]), l.filters)
                else:
                    cur = ((cur[0] - kh) // l.strides[0] + 1, (cur[1] - kw) // l.strides[1] + 1, l.filters)
            elif l.kind == "maxpool2d":
                cur = ((cur[0] - l.pool_size[0]) // l.strides[0] + 1, (cur[1] - l.pool_size[1]) // l.strides[1] + 1, cur[2])
            elif l.kind == "flatten":
                cur = (int(np.prod(cur)),)
        return cur


def _clone_layer(l: Layer) -> Layer:
    if l.kind == "dense":
        return Dense(l.units, activation=l.activation.name, name=l.name)
    if l.kind == "conv2d":
        return Conv2D(l.filters, l.kernel_size, l.strides, l.padding, l.activation.name, name=l.name)
    if l.kind == "maxpool2d":
        return MaxPooling2D(l.pool_size, l.strides, name=l.name)
    return Flatten(name=l.name)


def model_from_json(text: str) -> Sequential:
    """Rebuild a Sequential from Keras' ``to_json()`` output (the arch.json written by every
    ``save``: optimizer.py:274-276, bayesbybackprop.py:193-195, hmc.py:320-322)."""
    d = json.loads(text) if isinstance(text, str) else text
    if d.get("class_name") != "Sequential":
        raise ValueError("only Sequential models are supported (got %r)" % d.get("class_name"))
    cfg = d["config"]
    layer_cfgs = cfg["layers"] if isinstance(cfg, dict) else cfg
    m = Sequential(name=cfg.get("name", "sequential") if isinstance(cfg, dict) else "sequential")
    for lc in layer_cfgs:
        cls, c = lc["class_name"], lc["config"]
        bis = c.get("batch_input_shape") or c.get("batch_shape")
        if bis is not None and m.input_shape is None:
            m.input_shape = tuple(int(v) for v in bis[1:])
        if cls == "InputLayer":
            continue
        if cls == "Dense":
            l = Dense(c["units"], activation=c.get("activation"), use_bias=c.get("use_bias", True), name=c.get("name"))
        elif cls == "Conv2D":
            if c.get("data_format", "channels_last") != "channels_last":
                raise ValueError("Conv2D: only channels_last")
            l = Conv2D(c["filters"], c["kernel_size"], c.get("strides", (1, 1)), c.get("padding", "valid"),
                       c.get("activation"), use_bias=c.get("use_bias", True), name=c.get("name"))
        elif cls in ("MaxPooling2D", "MaxPool2D"):
            l = MaxPooling2D(c.get("pool_size", (2, 2)), c.get("strides"), c.get("padding", "valid"), name=c.get("name"))
        elif cls == "Flatten":
            l = Flatten(name=c.get("name"))
        elif cls in ("Dropout", "Activation") and cls == "Dropout":
            continue  # inference no-op
        else:
            raise ValueError("unsupported Keras layer %r" % cls)
        l._model = m
        m.layers.append(l)
    m.build()
    return m


def from_keras(model) -> Sequential:
    """Accept a real tf.keras Sequential (duck-typed: to_json + get_weights)."""
    if isinstance(model, Sequential):
        return model
    if not (hasattr(model, "to_json") and hasattr(model, "get_weights")):
        raise TypeError("expected a deepbayes_b200.Sequential or a Keras model, got %r" % type(model))
    m = model_from_json(model.to_json())
    m.set_weights(model.get_weights())
    return m
