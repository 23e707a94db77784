"""Minimal TensorFlow-API stand-in on PyTorch-CPU (TEST INFRASTRUCTURE, see ../README.md).

Only the calls made by the reference's hot-path sources are provided.  Each function restates the
documented TensorFlow semantics of the op it names.
"""
from __future__ import annotations

import contextlib
import inspect
import math as _pymath
import sys
import types
from typing import Any, Callable, List, Optional

import numpy as np
import torch
import torch.nn.functional as F

newaxis = None
float32 = torch.float32
float64 = torch.float64
int32 = torch.int32
int64 = torch.int64
uint8 = torch.uint8
bool = torch.bool  # noqa: A001


class TFShape(tuple):
    @property
    def rank(self):
        return len(self)

    def as_list(self):
        return list(self)


class Tensor(torch.Tensor):
    """torch.Tensor whose ``.shape`` behaves like ``tf.TensorShape`` (``.rank``, ``.as_list()``)."""

    @property
    def shape(self):  # type: ignore[override]
        return TFShape(super().shape)

    def numpy(self):  # type: ignore[override]
        return self.detach().cpu().as_subclass(torch.Tensor).numpy()

    def set_shape(self, shape):
        return None

    def __len__(self):
        return super().shape[0]

    # tf.Tensor is immutable: augmented assignment rebinds to a new tensor (and may broadcast)
    def __imul__(self, other):
        return self * other

    def __iadd__(self, other):
        return self + other

    def __isub__(self, other):
        return self - other

    def __itruediv__(self, other):
        return self / other

    def __getitem__(self, idx):
        """TF/NumPy-style indexing incl. reversed slices ``[::-1]`` (torch rejects negative steps)."""
        items = idx if isinstance(idx, tuple) else (idx,)
        flips, new_items, dim = [], [], 0
        for it in items:
            if it is None:
                new_items.append(it)
                continue
            if it is Ellipsis:
                new_items.append(it)
                dim = None
                continue
            if isinstance(it, slice) and it.step is not None and it.step < 0:
                assert it.step == -1 and it.start is None and it.stop is None and dim is not None
                flips.append(dim)
                new_items.append(slice(None))
            else:
                new_items.append(it)
            if dim is not None and not (isinstance(it, torch.Tensor) and it.dtype == torch.bool and it.dim() > 1):
                dim += 1
        base = torch.flip(self, dims=flips) if flips else self
        return torch.Tensor.__getitem__(base, tuple(new_items) if isinstance(idx, tuple) else new_items[0])

    def __bool__(self):
        return builtins_bool(self.detach().as_subclass(torch.Tensor).item())


import builtins as _b  # noqa: E402

builtins_bool = _b.bool


def _t(x, dtype=None) -> torch.Tensor:
    if isinstance(x, torch.Tensor):
        y = x if isinstance(x, Tensor) else x.as_subclass(Tensor)
    else:
        y = torch.as_tensor(np.ascontiguousarray(np.asarray(x))).as_subclass(Tensor)
        if y.dtype == torch.float64 and dtype is None:
            y = y.to(DEFAULT_FLOAT)
    if dtype is not None and y.dtype != dtype:
        y = y.to(dtype)
    return y


DEFAULT_FLOAT = torch.float32


# ----------------------------------------------------------------------------------------------
# variables
# ----------------------------------------------------------------------------------------------
class VariableSynchronization:
    ON_READ = "ON_READ"


class VariableAggregation:
    ONLY_FIRST_REPLICA = "ONLY_FIRST_REPLICA"


def Variable(initial_value, name=None, trainable=True, dtype=None, **_ignored):
    v = _t(initial_value, dtype).detach().clone().as_subclass(Tensor)
    if v.is_floating_point():
        v = v.to(DEFAULT_FLOAT).as_subclass(Tensor)
    v.requires_grad_(builtins_bool(trainable) and v.is_floating_point())
    v._tf_name = name or "Variable"
    v._tf_trainable = builtins_bool(trainable)
    v._tf_is_variable = True
    return v


def _assign(self, value):
    with torch.no_grad():
        self.as_subclass(torch.Tensor).copy_(torch.as_tensor(value).as_subclass(torch.Tensor))
    return self


Tensor.assign = _assign
Tensor.name = property(lambda self: getattr(self, "_tf_name", "tensor") + ":0")
Tensor.trainable = property(lambda self: getattr(self, "_tf_trainable", False))


# ----------------------------------------------------------------------------------------------
# random draws: served from a queue so that reference code and oracle consume identical tensors
# ----------------------------------------------------------------------------------------------
class _Draws:
    def __init__(self):
        self.queue: List[Any] = []
        self.gen = torch.Generator().manual_seed(0)
        self.log: List[tuple] = []

    def pop(self, kind, shape):
        if self.queue:
            k, v = self.queue.pop(0)
            assert k == kind, f"draw order mismatch: code asked for {kind}{tuple(shape)}, queue has {k}"
            if torch.is_tensor(v):
                assert tuple(v.shape) == tuple(shape), (kind, tuple(v.shape), tuple(shape))
            return v
        return None


DRAWS = _Draws()


class _Random(types.ModuleType):
    @staticmethod
    def normal(shape, mean=0.0, stddev=1.0, dtype=None, **_):
        shape = [int(s) for s in shape]
        v = DRAWS.pop("normal", shape)
        if v is None:
            v = torch.randn(*shape, generator=DRAWS.gen, dtype=torch.float64).to(DEFAULT_FLOAT)
        return _t(v.to(DEFAULT_FLOAT) * float(stddev) + float(mean))

    @staticmethod
    def uniform(shape, minval=0.0, maxval=1.0, dtype=None, **_):
        shape = [int(s) for s in shape]
        if dtype in (torch.int32, torch.int64):
            v = DRAWS.pop("uniform_int", shape)
            if v is None:
                v = torch.randint(int(minval), int(maxval), shape, generator=DRAWS.gen)
            return _t(torch.as_tensor(v), dtype)
        v = DRAWS.pop("uniform", shape)
        if v is None:
            v = torch.rand(*shape, generator=DRAWS.gen) * (maxval - minval) + minval
        return _t(torch.as_tensor(v, dtype=DEFAULT_FLOAT))


random = _Random("tensorflow.random")


# ----------------------------------------------------------------------------------------------
# array / math ops
# ----------------------------------------------------------------------------------------------
def shape(x):
    return TFShape(torch.Tensor.size(x)) if isinstance(x, torch.Tensor) else TFShape(np.shape(x))


def _ints(seq):
    return [int(s) for s in seq]


def reshape(x, shape):  # noqa: A002
    return _t(x).reshape(_ints(shape))


def transpose(x, perm=None):
    x = _t(x)
    return x.permute(*perm) if perm is not None else x.t()


def constant(value, dtype=None, shape=None):  # noqa: A002
    return _t(value, dtype)


def convert_to_tensor(value, dtype=None):
    return _t(value, dtype)


def identity(x):
    return _t(x) * 1 if isinstance(x, torch.Tensor) else _t(x)


def cast(x, dtype):
    if isinstance(x, (int, float)):
        return float(x) if dtype in (torch.float32, torch.float64) else int(x)
    return _t(x).to(DEFAULT_FLOAT if dtype == torch.float32 else dtype)


def zeros(shape, dtype=torch.float32):  # noqa: A002
    return _t(torch.zeros(_ints(shape), dtype=DEFAULT_FLOAT if dtype == torch.float32 else dtype))


def ones(shape, dtype=torch.float32):  # noqa: A002
    return _t(torch.ones(_ints(shape), dtype=DEFAULT_FLOAT if dtype == torch.float32 else dtype))


def ones_like(x, dtype=None):
    return _t(torch.ones_like(_t(x), dtype=(DEFAULT_FLOAT if dtype == torch.float32 else dtype)))


def range(start, limit=None, delta=1, dtype=None):  # noqa: A001
    if limit is None:
        start, limit = 0, start
    return _t(torch.arange(start, limit, delta), dtype)


def expand_dims(x, axis):
    return _t(x).unsqueeze(axis)


def tile(x, multiples):
    if isinstance(x, (list, tuple)) and not any(isinstance(e, torch.Tensor) for e in x):
        arr = np.tile(np.asarray(x), _ints(multiples) if not isinstance(multiples, int) else multiples)
        if arr.ndim == 1 and all(isinstance(e, int) for e in x):
            return [int(v) for v in arr]
        return _t(arr)
    if isinstance(x, (list, tuple)):
        x = torch.stack([_t(e) for e in x])
    return _t(x).repeat(*_ints(multiples))


def repeat(x, repeats, axis=None):
    reps = torch.as_tensor(repeats)
    return torch.repeat_interleave(_t(x), reps, dim=axis)


def concat(values, axis):
    return torch.cat([_t(v) for v in values], dim=axis)


def pad(x, paddings):
    flat = []
    for lo, hi in reversed([list(p) for p in paddings]):
        flat += [int(lo), int(hi)]
    return F.pad(_t(x), flat)


def where(condition, x=None, y=None):
    c = _t(condition)
    if x is None and y is None:
        return torch.nonzero(c)
    xx = x if isinstance(x, torch.Tensor) else torch.as_tensor(x, dtype=DEFAULT_FLOAT if isinstance(x, float) else None)
    yy = y if isinstance(y, torch.Tensor) else torch.as_tensor(y, dtype=DEFAULT_FLOAT if isinstance(y, float) else None)
    return _t(torch.where(c, xx, yy))


def broadcast_to(x, shape):  # noqa: A002
    return _t(x).expand(*_ints(shape))


def equal(a, b):
    return _t(a) == b


def reverse(x, axis):
    return torch.flip(_t(x), dims=list(axis))


def clip_by_value(x, lo, hi):
    return torch.clamp(_t(x), lo, hi)


def minimum(a, b):
    if isinstance(a, (int, float)) and isinstance(b, (int, float)):
        return min(a, b)
    return torch.minimum(_t(a), _t(b))


def maximum(a, b):
    if isinstance(a, (int, float)) and isinstance(b, (int, float)):
        return max(a, b)
    return torch.maximum(_t(a), _t(b))


def _axes(axis):
    if axis is None:
        return None
    return tuple(axis) if isinstance(axis, (list, tuple)) else (int(axis),)


def reduce_sum(x, axis=None, keepdims=False):
    x = _t(x)
    return x.sum() if axis is None else x.sum(dim=_axes(axis), keepdim=keepdims)


def reduce_mean(x, axis=None, keepdims=False):
    x = _t(x)
    return x.mean() if axis is None else x.mean(dim=_axes(axis), keepdim=keepdims)


def reduce_max(x, axis=None):
    x = _t(x)
    return x.max() if axis is None else x.amax(dim=_axes(axis))


def reduce_prod(x, axis=None):
    if not isinstance(x, torch.Tensor):
        return int(np.prod([int(v) for v in x]))
    return _t(x).prod()


def square(x):
    x = _t(x)
    return x * x


def sqrt(x):
    if isinstance(x, (int, float)):
        return _pymath.sqrt(x)
    return torch.sqrt(_t(x))


def matmul(a, b):
    return _t(a) @ _t(b)


def argmax(x, axis=None):
    return torch.argmax(_t(x), dim=axis)


def map_fn(fn, elems, dtype=None, **_):
    n = elems[0].shape[0] if isinstance(elems, (tuple, list)) else elems.shape[0]
    outs = []
    for i in builtins_range(n):
        outs.append(fn(tuple(e[i] for e in elems)) if isinstance(elems, (tuple, list)) else fn(elems[i]))
    return torch.stack(outs)


builtins_range = _b.range


class _Math(types.ModuleType):
    rsqrt = staticmethod(lambda x: torch.rsqrt(_t(x)) if isinstance(x, torch.Tensor) else 1.0 / _pymath.sqrt(x))
    sqrt = staticmethod(sqrt)
    square = staticmethod(square)
    softplus = staticmethod(lambda x: F.softplus(_t(x)))

    @staticmethod
    def floordiv(a, b):
        if isinstance(a, (int, float)) and isinstance(b, (int, float)):
            return a // b
        return torch.div(_t(a), b, rounding_mode="floor")


math_mod = _Math("tensorflow.math")


class _NN(types.ModuleType):
    @staticmethod
    def conv2d(x, filters, strides, padding, data_format="NHWC", **_):
        """tf.nn.conv2d: cross-correlation, filters HWIO; SAME pads (k-1)/2 for stride 1."""
        x, w = _t(x), _t(filters)
        if data_format == "NHWC":
            xin, s = x.permute(0, 3, 1, 2), (int(strides[1]), int(strides[2]))
        else:
            xin, s = x, (int(strides[2]), int(strides[3]))
        kh, kw = w.shape[0], w.shape[1]
        groups = xin.shape[1] // w.shape[2]
        wt = w.permute(3, 2, 0, 1)
        if padding == "SAME":
            assert s == (1, 1), "shim implements SAME only for stride 1 (all the reference needs)"
            y = F.conv2d(xin, wt, stride=s, padding=((kh - 1) // 2, (kw - 1) // 2), groups=groups)
        else:
            y = F.conv2d(xin, wt, stride=s, groups=groups)
        return y.permute(0, 2, 3, 1) if data_format == "NHWC" else y

    @staticmethod
    def conv2d_transpose(x, filters, output_shape, strides, padding="SAME", data_format="NHWC", **_):
        """tf.nn.conv2d_transpose (gradient of conv2d w.r.t. its input): filters [kh,kw,out_c,in_c]."""
        assert padding == "VALID"
        x, w = _t(x), _t(filters)
        if data_format == "NHWC":
            xin, s = x.permute(0, 3, 1, 2), (int(strides[1]), int(strides[2]))
        else:
            xin, s = x, (int(strides[2]), int(strides[3]))
        in_c = xin.shape[1]
        groups = in_c // w.shape[3] if w.shape[3] != in_c else 1
        y = F.conv_transpose2d(xin, w.permute(3, 2, 0, 1), stride=s, groups=groups)
        y = y.permute(0, 2, 3, 1) if data_format == "NHWC" else y
        assert [int(v) for v in output_shape[1:]] == list(y.shape[1:]), (output_shape, y.shape)
        return y

    @staticmethod
    def embedding_lookup(params, ids):
        return _t(params)[_t(ids).long()]

    @staticmethod
    def sparse_softmax_cross_entropy_with_logits(labels, logits):
        lg = _t(logits)
        out = F.cross_entropy(lg.reshape(-1, lg.shape[-1]), _t(labels).reshape(-1).long(), reduction="none")
        return out.reshape(lg.shape[:-1])


nn = _NN("tensorflow.nn")


class _Image(types.ModuleType):
    @staticmethod
    def resize(images, size, **_):
        """tf.image.resize default: bilinear, half-pixel centres, no antialias; [H,W,C] or [N,H,W,C]."""
        x = _t(images)
        if x.shape[1] == 0 or x.shape[-2] == 0:
            raise ValueError("resize of an empty image")
        if x.dim() == 3:
            return F.interpolate(x.permute(2, 0, 1)[None], size=_ints(size), mode="bilinear", align_corners=False)[0] \
                .permute(1, 2, 0)
        return F.interpolate(x.permute(0, 3, 1, 2), size=_ints(size), mode="bilinear", align_corners=False) \
            .permute(0, 2, 3, 1)


image = _Image("tensorflow.image")


# ----------------------------------------------------------------------------------------------
# autodiff, functions, distribution, misc
# ----------------------------------------------------------------------------------------------
class GradientTape:
    """torch tracks every op, so the tape only has to answer ``gradient`` (of any order)."""

    def __init__(self, persistent=False):
        self.persistent = persistent

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def watch(self, t):
        if isinstance(t, torch.Tensor) and not t.requires_grad:
            t.requires_grad_(True)

    def gradient(self, target, sources):
        single = isinstance(sources, torch.Tensor)
        src = [sources] if single else list(sources)
        grads = torch.autograd.grad(target, src, retain_graph=True, create_graph=True, allow_unused=True)
        grads = [None if g is None else g.as_subclass(Tensor) for g in grads]
        return grads[0] if single else grads


def function(fn=None, **_):
    if fn is None:
        return lambda f: f
    return fn


def custom_gradient(fn):
    raise NotImplementedError("custom_gradient (CUDA upfirdn path) is not reachable: is_built_with_cuda() is False")


class _Test(types.ModuleType):
    @staticmethod
    def is_built_with_cuda():
        return False


test = _Test("tensorflow.test")


class _Config(types.ModuleType):
    @staticmethod
    def list_physical_devices(kind=None):
        return []

    class experimental:
        @staticmethod
        def set_memory_growth(*a, **k):
            return None


config = _Config("tensorflow.config")


class _Data(types.ModuleType):
    class experimental:
        AUTOTUNE = -1


data = _Data("tensorflow.data")


class _Distribute(types.ModuleType):
    class ReduceOp:
        SUM = "SUM"
        MEAN = "MEAN"

    class Strategy:
        pass

    class MirroredStrategy(Strategy):
        num_replicas_in_sync = 1

        def run(self, fn, args=(), kwargs=None):
            return fn(*args, **(kwargs or {}))

        def reduce(self, reduce_op, value, axis=None):
            return value

        @contextlib.contextmanager
        def scope(self):
            yield self


distribute = _Distribute("tensorflow.distribute")


class _DTypes(types.ModuleType):
    float32 = torch.float32
    int32 = torch.int32
    uint8 = torch.uint8


dtypes = _DTypes("tensorflow.dtypes")
math = math_mod  # noqa: F811  (tf.math)

from . import keras  # noqa: E402,F401

for _name, _mod in (("random", random), ("nn", nn), ("math", math_mod), ("image", image), ("test", test),
                    ("config", config), ("data", data), ("distribute", distribute), ("dtypes", dtypes)):
    sys.modules[f"tensorflow.{_name}"] = _mod
