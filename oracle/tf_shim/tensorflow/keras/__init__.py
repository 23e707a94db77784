"""tf.keras subset: Layer / Model with build-on-first-call, variable tracking and Keras'
training-flag propagation; Dense, Dropout, ReLU, LeakyReLU, Lambda; optimizer_v2.Adam; losses.mse."""
from __future__ import annotations

import inspect
import math
import sys
import types

import torch
import torch.nn.functional as F

import tensorflow as tf

_TRAINING = [None]   # Keras call-context stack


def _collect(obj, seen, out, only_trainable):
    if id(obj) in seen:
        return
    seen.add(id(obj))
    if isinstance(obj, torch.Tensor) and getattr(obj, "_tf_is_variable", False):
        if not only_trainable or obj._tf_trainable:
            out.append(obj)
    elif isinstance(obj, Layer):
        for v in obj.__dict__.values():
            _collect(v, seen, out, only_trainable)
    elif isinstance(obj, (list, tuple)):
        for v in obj:
            _collect(v, seen, out, only_trainable)
    elif isinstance(obj, dict):
        for v in obj.values():
            _collect(v, seen, out, only_trainable)


class Layer:
    def __init__(self, name=None, **kwargs):
        self.name = name or type(self).__name__.lower()
        self.built = False

    def build(self, input_shape):
        pass

    def call(self, *args, **kwargs):
        raise NotImplementedError

    def get_config(self):
        return {}

    def _input_shape(self, x):
        if isinstance(x, torch.Tensor):
            return tf.TFShape(torch.Tensor.size(x))
        if isinstance(x, (list, tuple)):
            return [self._input_shape(e) for e in x]
        return None

    def __call__(self, *args, **kwargs):
        if not self.built:
            self.build(self._input_shape(args[0]) if args else None)
            self.built = True
        sig = inspect.signature(self.call).parameters
        has_training = "training" in sig
        # Keras 2.8: explicit value > call context > signature default > False
        if kwargs.get("training") is not None:
            val = kwargs["training"]
        elif _TRAINING[-1] is not None:
            val = _TRAINING[-1]
        elif has_training and sig["training"].default not in (inspect._empty, None):
            val = sig["training"].default
        else:
            val = False
        if has_training and kwargs.get("training") is None:
            bound_positionally = False
            names = list(sig.keys())
            if "training" in names and names.index("training") < len(args):
                bound_positionally = True
                val = args[names.index("training")]
            if not bound_positionally:
                kwargs["training"] = val
        _TRAINING.append(val)
        try:
            return self.call(*args, **kwargs)
        finally:
            _TRAINING.pop()

    @property
    def trainable_variables(self):
        out = []
        _collect(self, set(), out, True)
        return out

    @property
    def weights(self):
        out = []
        _collect(self, set(), out, False)
        return out

    def get_weights(self):
        return [w.detach().clone() for w in self.weights]

    def set_weights(self, values):
        for w, v in zip(self.weights, values):
            w.assign(v)


class Model(Layer):
    pass


class _Layers(types.ModuleType):
    Layer = Layer

    class Lambda(Layer):
        def __init__(self, fn, **kw):
            super().__init__(**kw)
            self.fn = fn

        def call(self, x):
            return self.fn(x)

    class LeakyReLU(Layer):
        def __init__(self, alpha=0.3, **kw):
            super().__init__(**kw)
            self.alpha = alpha

        def call(self, x):
            return F.leaky_relu(x, self.alpha)

    class ReLU(Layer):
        def call(self, x):
            return torch.relu(x)

    class Dropout(Layer):
        """Keras Dropout: at training time keep with prob 1-rate and scale kept values by 1/(1-rate)."""

        def __init__(self, rate, **kw):
            super().__init__(**kw)
            self.rate = rate

        def call(self, x, training=None):
            if not training:
                return x
            mask = tf.DRAWS.pop("dropout", list(x.shape))
            if mask is None:
                mask = (torch.rand(*x.shape, generator=tf.DRAWS.gen) >= self.rate).to(x.dtype)
            return x * mask.to(x.dtype) / (1.0 - self.rate)

    class Dense(Layer):
        """Keras Dense(units): glorot-uniform kernel, zero bias, linear activation."""

        def __init__(self, units, **kw):
            super().__init__(**kw)
            self.units = units

        def build(self, input_shape):
            fan_in = int(input_shape[-1])
            lim = math.sqrt(6.0 / (fan_in + self.units))
            k = (torch.rand(fan_in, self.units, generator=tf.DRAWS.gen, dtype=torch.float64) * 2 - 1) * lim
            self.kernel = tf.Variable(k, name="kernel")
            self.bias = tf.Variable(torch.zeros(self.units), name="bias")

        def call(self, x):
            return x @ self.kernel + self.bias


layers = _Layers("tensorflow.keras.layers")


class _Adam:
    """tf.keras.optimizers.Adam (optimizer_v2, non-amsgrad): epsilon outside the corrected sqrt."""

    class _It:
        def __init__(self):
            self.v = 0

        def numpy(self):
            return self.v

    def __init__(self, learning_rate=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-7, **_):
        self.lr, self.b1, self.b2, self.eps = learning_rate, beta_1, beta_2, epsilon
        self.iterations = self._It()
        self.m, self.v = {}, {}

    def apply_gradients(self, grads_and_vars):
        t = self.iterations.v + 1
        lr_t = self.lr * math.sqrt(1.0 - self.b2 ** t) / (1.0 - self.b1 ** t)
        for g, var in grads_and_vars:
            if g is None:
                continue
            g = g.detach().as_subclass(torch.Tensor)
            k = id(var)
            if k not in self.m:
                self.m[k] = torch.zeros_like(g)
                self.v[k] = torch.zeros_like(g)
            self.m[k] = self.b1 * self.m[k] + (1 - self.b1) * g
            self.v[k] = self.b2 * self.v[k] + (1 - self.b2) * g * g
            var.assign(var.detach().as_subclass(torch.Tensor) - lr_t * self.m[k] / (torch.sqrt(self.v[k]) + self.eps))
        self.iterations.v = t


class _Optimizers(types.ModuleType):
    Adam = _Adam


optimizers = _Optimizers("tensorflow.keras.optimizers")


class _Losses(types.ModuleType):
    @staticmethod
    def mse(y_true, y_pred):
        return ((y_pred - y_true) ** 2).mean(dim=-1)


losses = _Losses("tensorflow.keras.losses")

from . import preprocessing  # noqa: E402,F401

sys.modules["tensorflow.keras.layers"] = layers
sys.modules["tensorflow.keras.optimizers"] = optimizers
sys.modules["tensorflow.keras.losses"] = losses


class _Metrics(types.ModuleType):
    class Mean:  # utils/loss_tracker.py imports it; not on the hot path
        def __init__(self, name=None):
            self.total, self.count = 0.0, 0

        def __call__(self, v):
            self.total += float(v)
            self.count += 1

        def result(self):
            return self.total / max(self.count, 1)

        def reset_states(self):
            self.total, self.count = 0.0, 0


metrics = _Metrics("tensorflow.keras.metrics")
sys.modules["tensorflow.keras.metrics"] = metrics
