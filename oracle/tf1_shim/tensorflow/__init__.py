"""Eager stand-in for the TensorFlow 1.8 symbols that AlignNet-3D's hot path touches.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  TensorFlow 1.8 -- the third-party, un-vendored
dependency in which the reference's arithmetic lives (README.md:31) -- cannot be installed in this
image (Python 3.12, no wheel, no network).  This package lets the reference's OWN, UNMODIFIED
`models/tp8.py` and `utils/tf_util.py` be imported from /root/reference and executed eagerly, so
that golden vectors come from the reference's call sites (scopes, op order, slicing, broadcasting
quirks, loss assembly) instead of from a restatement of them.  Only the primitive ops are restated
here, each following TensorFlow 1.8's published definition:

  tf.nn.conv2d            NHWC x HWIO cross-correlation        -> torch.nn.functional.conv2d
  tf.nn.moments           mean; variance = mean((x-stop_gradient(mean))^2)   (biased)
  tf.nn.batch_normalization   x*(rsqrt(var+eps)*scale) + (offset - mean*rsqrt(var+eps)*scale)
  tf.nn.max_pool          NHWC window max                      -> torch.nn.functional.max_pool2d
  tf.nn.dropout           x / keep_prob * bernoulli(keep_prob)   (masks can be injected)
  tf.train.ExponentialMovingAverage   zero-initialised shadow of a Tensor; shadow -= (1-decay)*(shadow-value)
  tf.mod                  floor-mod (sign of the divisor)      tf.argmax: first maximal index
  tf.variable_scope / tf.get_variable / tf.Variable   TF1 naming: get_variable is keyed by the
      VARIABLE scope (shared under reuse), tf.Variable by the uniquified NAME scope (`siamese_1/...`
      on the second siamese pass), `variable_scope('')` adds no prefix.

Tensors are torch tensors (autograd gives `tf.gradients`).  `set_float_dtype(torch.float64)` runs the
same graph in double precision for a rounding-free ground truth.  Variables live in a module-level
store keyed by their TF names; `preload()` injects initial values, `reset()` clears everything.
"""
from __future__ import annotations

import builtins
import contextlib
import types
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn.functional as F

__version__ = "1.8.0-shim"

# ------------------------------------------------------------------------------------------------
# dtypes
# ------------------------------------------------------------------------------------------------
_FLOAT = torch.float32


class _DType:
    def __init__(self, name):
        self.name = name

    @property
    def torch(self):
        return {"float32": _FLOAT, "float16": torch.float16, "float64": torch.float64, "int32": torch.int32,
                "int64": torch.int64, "bool": torch.bool}[self.name]

    def __repr__(self):
        return "tf." + self.name


float32, float16, float64 = _DType("float32"), _DType("float16"), _DType("float64")
int32, int64 = _DType("int32"), _DType("int64")
AUTO_REUSE = "AUTO_REUSE"


def set_float_dtype(dtype) -> None:
    """tf.float32 tensors are created in `dtype` (torch.float32 = TF semantics, torch.float64 = exact)."""
    global _FLOAT
    _FLOAT = dtype


def _td(dtype):
    if dtype is None:
        return None
    if isinstance(dtype, _DType):
        return dtype.torch
    return dtype


# ------------------------------------------------------------------------------------------------
# tensors with the TF1 static-shape API
# ------------------------------------------------------------------------------------------------
class Dimension(builtins.int):
    @property
    def value(self):
        return builtins.int(self)


class TensorShape(list):
    def as_list(self):
        return [builtins.int(d) for d in self]


class Tensor(torch.Tensor):
    """torch.Tensor + `get_shape()` (models/tp8.py:102,137; utils/tf_util.py:143,326,468)."""

    def get_shape(self):
        return TensorShape(Dimension(d) for d in self.shape)


def _wrap(t) -> Tensor:
    return t if isinstance(t, Tensor) else t.as_subclass(Tensor)


def _t(x, like=None, dtype=None) -> torch.Tensor:
    """Python number / numpy / tensor -> tensor (floats in the active float dtype)."""
    if isinstance(x, torch.Tensor):
        return x
    if dtype is None:
        if like is not None and isinstance(like, torch.Tensor) and not isinstance(x, (builtins.bool,)):
            dtype = like.dtype if (like.is_floating_point() or isinstance(x, builtins.int)) else _FLOAT
        else:
            arr = np.asarray(x)
            dtype = _FLOAT if arr.dtype.kind == "f" else (torch.int32 if arr.dtype.kind in "iu" else torch.bool)
    return _wrap(torch.as_tensor(np.asarray(x), dtype=dtype))


# ------------------------------------------------------------------------------------------------
# scopes, variables
# ------------------------------------------------------------------------------------------------
class _State:
    def __init__(self):
        self.reset()

    def reset(self):
        self.var_scope: List[str] = []          # variable-scope path (get_variable names)
        self.reuse = None
        self.name_scope: str = ""               # current name scope, '' or 'a/b' (tf.Variable names)
        self.used_names: Dict[str, int] = {}    # name uniquification, per graph
        self.variables: Dict[str, Tensor] = {}  # created variables by full name
        self.trainable: List[str] = []
        self.shadows: Dict[str, Tensor] = {}    # EMA shadow variables by full name
        self.preloaded: Dict[str, np.ndarray] = {}
        self.dropout_masks: List[np.ndarray] = []
        self.dropout_calls = 0


_S = _State()


def reset() -> None:
    _S.reset()


def preload(values: Dict[str, np.ndarray]) -> None:
    """Initial values by TF variable name (overrides initialisers; EMA shadows included)."""
    _S.preloaded.update(values)


def queue_dropout_masks(masks) -> None:
    """Keep-masks consumed by successive tf.nn.dropout calls, in graph-construction order."""
    _S.dropout_masks = list(masks)
    _S.dropout_calls = 0


def variables() -> Dict[str, Tensor]:
    return dict(_S.variables)


def trainable_variables() -> List[str]:
    return list(_S.trainable)


def shadow_variables() -> Dict[str, Tensor]:
    return dict(_S.shadows)


def _unique(name: str) -> str:
    """TF graph.unique_name: first use keeps the name, later uses get _1, _2, ..."""
    n = _S.used_names.get(name, 0)
    _S.used_names[name] = n + 1
    if n == 0:
        return name
    cand = f"{name}_{n}"
    while cand in _S.used_names:
        n += 1
        cand = f"{name}_{n}"
    _S.used_names[cand] = 1
    return cand


class _VarScope:
    def __init__(self, name):
        self.name = name


@contextlib.contextmanager
def variable_scope(name_or_scope, reuse=None, **_):
    name = name_or_scope if isinstance(name_or_scope, str) else name_or_scope.name.split("/")[-1]
    old = (list(_S.var_scope), _S.reuse, _S.name_scope)
    if name:
        _S.var_scope.append(name)
        full = (_S.name_scope + "/" if _S.name_scope else "") + name
        _S.name_scope = _unique(full)
    if reuse is not None:
        _S.reuse = reuse                          # inherited by nested scopes, as in TF
    try:
        yield _VarScope("/".join(_S.var_scope))
    finally:
        _S.var_scope, _S.reuse, _S.name_scope = old


@contextlib.contextmanager
def name_scope(name, *_, **__):
    old = _S.name_scope
    if name:
        _S.name_scope = _unique((_S.name_scope + "/" if _S.name_scope else "") + name)
    else:
        _S.name_scope = ""
    try:
        yield _S.name_scope
    finally:
        _S.name_scope = old


@contextlib.contextmanager
def device(_):
    yield


@contextlib.contextmanager
def control_dependencies(_):
    yield


def _new_variable(full: str, init: torch.Tensor, trainable: bool) -> Tensor:
    if full in _S.preloaded:
        pre = np.asarray(_S.preloaded[full])
        if tuple(pre.shape) != tuple(init.shape):
            pre = pre.reshape(tuple(init.shape))   # e.g. [Cin,Cout] matrices for [kh,kw,Cin,Cout] kernels
        init = torch.as_tensor(pre, dtype=init.dtype)
    v = _wrap(init.detach().clone())
    if trainable and v.is_floating_point():
        v.requires_grad_(True)
        _S.trainable.append(full)
    _S.variables[full] = v
    return v


def get_variable(name, shape=None, initializer=None, dtype=None, trainable=True, **_):
    full = "/".join(_S.var_scope + [name])
    if full in _S.variables:
        if _S.reuse in (True, AUTO_REUSE):
            return _S.variables[full]
        raise ValueError(f"Variable {full} already exists, disallowed. Did you mean to set reuse=True?")
    if _S.reuse is True:
        raise ValueError(f"Variable {full} does not exist, or was not created with tf.get_variable().")
    shape = [builtins.int(s) for s in shape]
    init = initializer(shape, _td(dtype) or _FLOAT) if callable(initializer) else _t(initializer).reshape(shape)
    return _new_variable(full, init, trainable)


def Variable(initial_value, name=None, trainable=True, **_):
    full = _unique((_S.name_scope + "/" if _S.name_scope else "") + (name or "Variable"))
    init = _t(initial_value)
    return _new_variable(full, init, trainable)


def constant_initializer(value=0.0):
    return lambda shape, dtype: torch.full(shape, builtins.float(value), dtype=dtype)


def truncated_normal_initializer(mean=0.0, stddev=1.0, seed=None):
    def init(shape, dtype):
        t = torch.empty(shape, dtype=dtype)
        torch.nn.init.trunc_normal_(t, mean, stddev, mean - 2 * stddev, mean + 2 * stddev)
        return t
    return init


def _xavier_initializer(uniform=True, seed=None, dtype=None):
    """tf.contrib.layers.xavier_initializer: U(-l, l), l = sqrt(6/(fan_in+fan_out)), fans include the window."""
    def init(shape, dtype):
        receptive = builtins.int(np.prod(shape[:-2])) if len(shape) > 2 else 1
        fan_in, fan_out = shape[-2] * receptive, shape[-1] * receptive
        limit = float(np.sqrt(6.0 / (fan_in + fan_out)))
        return torch.empty(shape, dtype=dtype).uniform_(-limit, limit)
    return init


contrib = types.SimpleNamespace(layers=types.SimpleNamespace(xavier_initializer=_xavier_initializer))


# ------------------------------------------------------------------------------------------------
# graph plumbing that is a no-op in eager execution
# ------------------------------------------------------------------------------------------------
class Graph:
    @contextlib.contextmanager
    def as_default(self):
        yield self


def placeholder(dtype, shape=None, name=None):
    return _wrap(torch.zeros([builtins.int(s) for s in (shape or [])], dtype=_td(dtype)))


def add_to_collection(*_, **__):
    return None


def no_op(*_, **__):
    return None


summary = types.SimpleNamespace(scalar=lambda *a, **k: None, histogram=lambda *a, **k: None,
                                merge_all=lambda *a, **k: None)


def identity(x, name=None):
    return x


def cond(pred, true_fn=None, false_fn=None, **kw):
    true_fn = true_fn or kw.get("fn1")
    false_fn = false_fn or kw.get("fn2")
    return true_fn() if builtins.bool(pred) else false_fn()


def map_fn(fn, elems, dtype=None, **_):
    return stack([fn(e) for e in elems])


# ------------------------------------------------------------------------------------------------
# element-wise and shape ops
# ------------------------------------------------------------------------------------------------
def constant(value, dtype=None, shape=None, name=None):
    t = _t(value, dtype=_td(dtype))
    if shape is not None:
        t = _wrap(torch.broadcast_to(t, [builtins.int(s) for s in shape]).clone())
    return t


def zeros(shape, dtype=float32, name=None):
    return _wrap(torch.zeros([builtins.int(s) for s in shape], dtype=_td(dtype)))


def cast(x, dtype, name=None):
    return _wrap(_t(x).to(_td(dtype)))


def to_float(x, name=None):
    return cast(x, float32)


def to_int32(x, name=None):
    return cast(x, int32)          # float -> int32 truncates toward zero, like tf.cast


def to_int64(x, name=None):
    return cast(x, int64)


def expand_dims(x, axis, name=None):
    return _wrap(torch.unsqueeze(_t(x), axis))


def tile(x, multiples, name=None):
    return _wrap(_t(x).repeat(*[builtins.int(m) for m in multiples]))


def reshape(x, shape, name=None):
    return _wrap(torch.reshape(_t(x), [builtins.int(s) for s in shape]))


def stack(values, axis=0, name=None):
    like = next((v for v in values if isinstance(v, torch.Tensor)), None)
    return _wrap(torch.stack([_t(v, like=like) for v in values], dim=axis))


def concat(values, axis, name=None):
    return _wrap(torch.cat(list(values), dim=axis))


def transpose(x, perm=None, name=None):
    x = _t(x)
    if perm is None:
        perm = list(builtins.range(x.dim()))[::-1]
    return _wrap(x.permute(*perm))


def range(*args, dtype=None, **_):   # noqa: A001  (TF name)
    args = [builtins.int(a) for a in args]
    return _wrap(torch.arange(*args, dtype=_td(dtype) or torch.int32))


def matmul(a, b, name=None):
    return _wrap(torch.matmul(a, b))


def multiply(a, b, name=None):
    return _wrap(_t(a, like=b if isinstance(b, torch.Tensor) else None) * _t(b, like=a if isinstance(a, torch.Tensor) else None))


def square(x, name=None):
    return _wrap(_t(x) * _t(x))


def abs(x, name=None):   # noqa: A001
    return _wrap(torch.abs(_t(x)))


def cos(x, name=None):
    return _wrap(torch.cos(_t(x)))


def sin(x, name=None):
    return _wrap(torch.sin(_t(x)))


def acos(x, name=None):
    return _wrap(torch.acos(_t(x)))


def minimum(a, b, name=None):
    a = _t(a, like=b if isinstance(b, torch.Tensor) else None)
    return _wrap(torch.minimum(a, _t(b, like=a).to(a.dtype)))


def maximum(a, b, name=None):
    a = _t(a, like=b if isinstance(b, torch.Tensor) else None)
    return _wrap(torch.maximum(a, _t(b, like=a).to(a.dtype)))


def mod(a, b, name=None):
    """tf.mod == floormod: the result has the sign of the divisor."""
    a = _t(a, like=b if isinstance(b, torch.Tensor) else None)
    return _wrap(torch.remainder(a, _t(b, like=a)))


def where(condition, x=None, y=None, name=None):
    return _wrap(torch.where(condition, _t(x, like=y if isinstance(y, torch.Tensor) else None),
                             _t(y, like=x if isinstance(x, torch.Tensor) else None)))


def _axes(axis):
    if axis is None:
        return None
    return tuple(axis) if isinstance(axis, (list, tuple)) else (axis,)


def reduce_mean(x, axis=None, keepdims=False, name=None, **kw):
    x = _t(x)
    ax = _axes(axis if axis is not None else kw.get("reduction_indices"))
    return _wrap(x.mean() if ax is None else x.mean(dim=ax, keepdim=keepdims))


def reduce_sum(x, axis=None, keepdims=False, name=None, **kw):
    x = _t(x)
    ax = _axes(axis if axis is not None else kw.get("reduction_indices"))
    return _wrap(x.sum() if ax is None else x.sum(dim=ax, keepdim=keepdims))


def reduce_max(x, axis=None, keepdims=False, name=None, **kw):
    x = _t(x)
    ax = _axes(axis)
    return _wrap(x.max() if ax is None else torch.amax(x, dim=ax, keepdim=keepdims))


def norm(x, ord="euclidean", axis=None, keepdims=False, name=None):
    return _wrap(torch.sqrt(torch.sum(_t(x) * _t(x), dim=axis, keepdim=keepdims)))


def argmax(x, axis=None, output_type=int64, name=None, **kw):
    """First maximal index along `axis` (torch.argmax has the same tie rule)."""
    ax = axis if axis is not None else kw.get("dimension", 0)
    return _wrap(torch.argmax(_t(x), dim=ax).to(_td(output_type)))


def gather_nd(params, indices, name=None):
    idx = indices.long()
    return _wrap(params[tuple(idx[..., i] for i in builtins.range(idx.shape[-1]))])


def one_hot(indices, depth, on_value=None, off_value=None, axis=-1, dtype=None, name=None):
    assert axis == -1
    oh = F.one_hot(indices.long(), builtins.int(depth))
    if dtype is not None:
        dt = _td(dtype)
    elif on_value is not None:
        dt = _FLOAT if isinstance(on_value, builtins.float) else torch.int32
    else:
        dt = _FLOAT
    on = 1 if on_value is None else on_value
    off = 0 if off_value is None else off_value
    return _wrap(torch.where(oh.bool(), torch.as_tensor(on, dtype=dt), torch.as_tensor(off, dtype=dt)))


manip = types.SimpleNamespace(roll=lambda x, shift, axis: _wrap(torch.roll(x, shift, dims=axis)))


class _Normal:
    def __init__(self, loc, scale):
        self.loc = stack(loc) if isinstance(loc, (list, tuple)) else _t(loc)
        self.scale = _t(scale, like=self.loc)

    def cdf(self, x):
        return _wrap(0.5 * (1.0 + torch.erf((x - self.loc) / (self.scale * np.sqrt(2.0)))))


distributions = types.SimpleNamespace(Normal=_Normal)


# ------------------------------------------------------------------------------------------------
# tf.nn
# ------------------------------------------------------------------------------------------------
def _conv2d(input, filter, strides, padding, **_):   # noqa: A002
    """NHWC input, HWIO filter, cross-correlation.  Only VALID padding occurs on the path."""
    assert padding == "VALID", "shim implements the VALID convolutions of the tp8 path"
    y = F.conv2d(input.permute(0, 3, 1, 2), filter.permute(3, 2, 0, 1), stride=(strides[1], strides[2]))
    return _wrap(y.permute(0, 2, 3, 1))


def _bias_add(value, bias, **_):
    return _wrap(value + bias)


def _max_pool(value, ksize, strides, padding, name=None, **_):
    assert padding == "VALID"
    y = F.max_pool2d(value.permute(0, 3, 1, 2), kernel_size=(ksize[1], ksize[2]), stride=(strides[1], strides[2]))
    return _wrap(y.permute(0, 2, 3, 1))


def _avg_pool(value, ksize, strides, padding, name=None, **_):
    assert padding == "VALID"
    y = F.avg_pool2d(value.permute(0, 3, 1, 2), kernel_size=(ksize[1], ksize[2]), stride=(strides[1], strides[2]))
    return _wrap(y.permute(0, 2, 3, 1))


def _moments(x, axes, name=None, keep_dims=False, **_):
    """TF 1.8 nn_impl.moments: mean, then mean of squared_difference(x, stop_gradient(mean)); squeezed."""
    ax = tuple(axes)
    mean = x.mean(dim=ax, keepdim=True)
    var = ((x - mean.detach()) ** 2).mean(dim=ax, keepdim=True)
    if not keep_dims:
        mean, var = mean.squeeze(ax), var.squeeze(ax)
    mean, var = _wrap(mean), _wrap(var)
    scope = (_S.name_scope + "/" if _S.name_scope else "") + (name or "moments")
    mean._shim_name, var._shim_name = scope + "/Squeeze", scope + "/Squeeze_1"
    return mean, var


def _batch_normalization(x, mean, variance, offset, scale, variance_epsilon, name=None):
    inv = torch.rsqrt(variance + variance_epsilon)
    if scale is not None:
        inv = inv * scale
    return _wrap(x * inv + ((offset - mean * inv) if offset is not None else (-mean * inv)))


def _dropout(x, keep_prob, noise_shape=None, seed=None, name=None):
    i = _S.dropout_calls
    _S.dropout_calls += 1
    if i < len(_S.dropout_masks):
        keep = torch.as_tensor(np.asarray(_S.dropout_masks[i]), dtype=x.dtype)
    else:
        shape = list(x.shape) if noise_shape is None else [builtins.int(s) for s in noise_shape]
        keep = torch.floor(torch.rand(shape, dtype=x.dtype) + keep_prob)
    return _wrap(x / keep_prob * keep)


def _sparse_ce(_sentinel=None, labels=None, logits=None, name=None):
    return _wrap(F.cross_entropy(logits, labels.long(), reduction="none"))


def _soft_ce(_sentinel=None, labels=None, logits=None, dim=-1, name=None):
    return _wrap(-(labels * F.log_softmax(logits, dim=dim)).sum(dim=dim))


nn = types.SimpleNamespace(
    conv2d=_conv2d, bias_add=_bias_add, relu=lambda x, name=None: _wrap(torch.relu(x)), max_pool=_max_pool,
    avg_pool=_avg_pool, moments=_moments, batch_normalization=_batch_normalization, dropout=_dropout,
    sparse_softmax_cross_entropy_with_logits=_sparse_ce, softmax_cross_entropy_with_logits_v2=_soft_ce,
    l2_loss=lambda t, name=None: _wrap((t * t).sum() / 2), softmax=lambda x, axis=-1, name=None: _wrap(F.softmax(x, dim=axis)))


# ------------------------------------------------------------------------------------------------
# tf.train
# ------------------------------------------------------------------------------------------------
class _EMA:
    """tf.train.ExponentialMovingAverage over Tensors: apply() creates a zero-initialised shadow named
    '<tensor op name>/ExponentialMovingAverage' and performs shadow -= (1 - decay) * (shadow - value)."""

    def __init__(self, decay, num_updates=None, zero_debias=False, name="ExponentialMovingAverage"):
        self.decay, self.name = decay, name

    def _shadow(self, t):
        key = t._shim_name + "/" + self.name
        if key not in _S.shadows:
            init = torch.zeros_like(t.detach())
            if key in _S.preloaded:
                init = torch.as_tensor(np.asarray(_S.preloaded[key]), dtype=init.dtype).reshape(init.shape)
            _S.shadows[key] = _wrap(init.clone())
        return _S.shadows[key]

    def apply(self, var_list=None):
        with torch.no_grad():
            for t in var_list:
                s = self._shadow(t)
                d = self.decay if isinstance(self.decay, torch.Tensor) else torch.as_tensor(self.decay, dtype=s.dtype)
                s -= (1.0 - d.to(s.dtype)) * (s - t.detach())
        return None

    def average(self, t):
        return self._shadow(t)


def _exponential_decay(learning_rate, global_step, decay_steps, decay_rate, staircase=False, name=None):
    p = _t(global_step).to(torch.float64) / builtins.float(decay_steps)
    if staircase:
        p = torch.floor(p)
    return _wrap((builtins.float(learning_rate) * builtins.float(decay_rate) ** p).to(_FLOAT))


train = types.SimpleNamespace(ExponentialMovingAverage=_EMA, exponential_decay=_exponential_decay)

bool = _DType("bool")   # noqa: A001  (last: the functions above use builtins.bool)
