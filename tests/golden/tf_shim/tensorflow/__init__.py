"""A minimal stand-in for the `tensorflow` 2.3 API surface MiladShahidi/BERT4ClickPath uses.

TEST INFRASTRUCTURE, never a product path.  TensorFlow 2.3.1 cannot be installed in the build
container, so the reference's own source files (imported UNMODIFIED from /root/reference by
tests/golden/make_reference_golden.py) are executed on top of this module instead: the
reference supplies the algorithm - which ops, in which order, on which tensors - and this file
supplies the published meaning of each op, eagerly, on torch CPU tensors (torch autograd stands
in for tf.GradientTape).  Every function names the TensorFlow behaviour it reproduces.

Float precision: `tf.float32` is torch.float32, or torch.float64 when TFSHIM_FLOAT64=1 (the
same graph in double precision - the "truth" the NumPy oracle's float64 mode is compared with;
constants cast to tf.float32 from NumPy / Python values are still rounded to float32 first, as
their TensorFlow values would be).

String tensors are NumPy object arrays; numeric tensors are `Tensor`, a torch.Tensor subclass
whose augmented assignments rebind instead of mutating (TensorFlow tensors are immutable).
Nothing here is imported by bert4clickpath_b200/ or by the tests that run on the GPU box: the
outputs are frozen in tests/golden/reference_*.npz.
"""
import builtins as _b
import math as _math
import os as _os
import sys as _sys
import types as _types

import numpy as _np
import torch as _torch

__version__ = "2.3.1-shim"
_F64 = _os.environ.get("TFSHIM_FLOAT64") == "1"

float32 = _torch.float64 if _F64 else _torch.float32
float64 = _torch.float64
int32 = _torch.int32
int64 = _torch.int64
string = _np.dtype(object)
newaxis = None
_bool = _torch.bool


# ================================================================================== tensors
class Tensor(_torch.Tensor):
    """Immutable-by-convention numeric tensor."""

    def __iadd__(self, other):
        return self + other

    def __isub__(self, other):
        return self - other

    def __imul__(self, other):
        return self * other

    def __itruediv__(self, other):
        return self / other

    def numpy(self):
        return self.detach().as_subclass(_torch.Tensor).cpu().numpy()

    __hash__ = _torch.Tensor.__hash__


def _is_str_dtype(dt):
    return isinstance(dt, _np.dtype) and dt.kind in "OUS"


def _is_str(x):
    if isinstance(x, (str, bytes)):
        return True
    if isinstance(x, _np.ndarray):
        return x.dtype.kind in "OUS"
    if isinstance(x, (list, tuple)) and len(x) > 0:
        return _is_str(x[0])
    return False


def _s(x):
    """A string tensor: NumPy object array of Python str."""
    if isinstance(x, _np.ndarray) and x.dtype == object:
        return x
    a = _np.empty(_np.shape(x), dtype=object)
    a[...] = _np.asarray(x, dtype=object) if not isinstance(x, (str, bytes)) else x
    return a


def _round_f32(t):
    """float64 mode: values that enter the graph as float32 constants keep float32 values."""
    return t.to(_torch.float32).to(float32)


def _t(x, dtype=None):
    """tf.convert_to_tensor: Python floats -> float32, Python ints -> int32, bool -> bool."""
    if isinstance(x, Variable):
        x = x.value
    if isinstance(x, _torch.Tensor):
        t = x if isinstance(x, Tensor) else x.as_subclass(Tensor)
        return t if dtype is None or t.dtype == dtype else t.to(dtype)
    if isinstance(x, _Shape):
        x = list(x)
    if isinstance(x, _np.ndarray):
        t = _torch.from_numpy(_np.ascontiguousarray(x)).clone()
        if t.dtype == _torch.float64 and dtype == float32:
            t = _round_f32(t)
        elif t.dtype == _torch.float32:
            t = t.to(float32)      # a NumPy float32 array is a tf.float32 tensor (float64 mode: widened)
    elif isinstance(x, _b.range):
        t = _torch.tensor(list(x), dtype=_torch.int32)
    else:
        t = _torch.as_tensor(x)
        if t.dtype == _torch.float64 or (t.dtype == _torch.float32 and not isinstance(x, _np.generic)):
            t = _round_f32(t.to(_torch.float64)) if _F64 else t.to(_torch.float32)
        elif t.dtype == _torch.int64 and not isinstance(x, _np.generic):
            t = t.to(_torch.int32)
    t = t.as_subclass(Tensor)
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t


def _like(x, ref):
    """A Python scalar takes the dtype of the tensor it meets (TensorFlow's scalar conversion)."""
    if isinstance(x, (_b.int, _b.float, _b.bool)) and not isinstance(ref, (_b.int, _b.float, _b.bool)):
        return _t(x, _t(ref).dtype)
    return _t(x)


class _Shape(tuple):
    """tf.shape(x): usable as a Python tuple of ints (eager mode)."""


class Variable:
    """A scalar / tensor resource variable (metric accumulators; layer weights use Parameter)."""

    def __init__(self, initial_value, name=None, trainable=False, dtype=None):
        self.value = _t(initial_value, dtype)
        self.name = name

    def assign(self, v):
        self.value = _t(v, self.value.dtype).reshape(self.value.shape).clone()
        return self

    def assign_add(self, v):
        self.value = self.value + _t(v, self.value.dtype)
        return self

    def numpy(self):
        return self.value.numpy()

    def __truediv__(self, o):
        return self.value / _t(o)

    def __rtruediv__(self, o):
        return _like(o, self.value) / self.value

    def __mul__(self, o):
        return self.value * _like(o, self.value)

    __rmul__ = __mul__

    def __add__(self, o):
        return self.value + _like(o, self.value)

    __radd__ = __add__

    def __sub__(self, o):
        return self.value - _like(o, self.value)


class TensorSpec:
    def __init__(self, shape=None, dtype=None, name=None):
        self.shape, self.dtype, self.name = shape, dtype, name


class TensorShape(list):
    pass


class SparseTensor:
    def __init__(self, indices, values, dense_shape):
        self.indices, self.values, self.dense_shape = indices, values, dense_shape


class RaggedTensor:
    """Rows of a flat value list (tf.RaggedTensor with row-id partitioning)."""

    def __init__(self, values, value_rowids, nrows):
        self.values, self.value_rowids, self.nrows = values, value_rowids, _b.int(nrows)

    @classmethod
    def from_value_rowids(cls, values, value_rowids, nrows=None):
        rid = _t(value_rowids).to(_torch.int64)
        if nrows is None:
            nrows = _b.int(rid.max()) + 1 if rid.numel() else 0
        assert _b.bool((rid[1:] >= rid[:-1]).all()), "value_rowids must be sorted"
        return cls(values, rid, nrows)

    def to_tensor(self, default_value=0):
        """Right-pad every row to the longest one (tf.RaggedTensor.to_tensor)."""
        v = _t(self.values)
        rid = self.value_rowids
        counts = _torch.bincount(rid, minlength=self.nrows) if rid.numel() else _torch.zeros(
            self.nrows, dtype=_torch.int64)
        width = _b.int(counts.max()) if self.nrows > 0 and rid.numel() else 0
        starts = _torch.cumsum(counts, 0) - counts
        pos = _torch.arange(rid.numel()) - starts[rid]
        out = _torch.full((self.nrows, width) + tuple(v.shape[1:]), default_value, dtype=v.dtype)
        out = _torch.index_put(out, (rid, pos), v.as_subclass(_torch.Tensor))
        return out.as_subclass(Tensor)


# ====================================================================================== ops
def _cast(x, dtype):
    if _is_str_dtype(dtype) if isinstance(dtype, _np.dtype) else False:
        return _s(x)
    if isinstance(x, _np.ndarray) and x.dtype == _np.float64 and dtype == float32:
        return _round_f32(_torch.from_numpy(_np.ascontiguousarray(x))).as_subclass(Tensor)
    return _t(x, dtype)


def _shape(x):
    if _is_str(x) or isinstance(x, _np.ndarray):
        return _Shape(_np.shape(x))
    return _Shape(_t(x).shape)


def _size(x):
    return _b.int(_np.size(x)) if isinstance(x, _np.ndarray) else _b.int(_t(x).numel())


def _equal(a, b):
    if _is_str(a) or _is_str(b):
        return _torch.from_numpy(_np.asarray(_s(a) == _s(b) if not isinstance(b, str) else _s(a) == b,
                                             dtype=_np.bool_)).as_subclass(Tensor)
    a = _like(a, b)
    return a == _like(b, a)


def _where(condition, x=None, y=None):
    c = _t(condition).to(_torch.bool)
    if x is None and y is None:
        return _torch.nonzero(c).as_subclass(Tensor)          # (n, rank) int64, row-major order
    xt = _like(x, y) if isinstance(x, (_b.int, _b.float)) and not isinstance(x, _b.bool) else _t(x)
    yt = _like(y, xt)
    return _torch.where(c, xt, yt)


def _reduce(fn):
    def op(input_tensor, axis=None, keepdims=False):
        t = _t(input_tensor)
        if t.dtype == _torch.bool:
            t = t.to(_torch.int32)
        return fn(t) if axis is None else fn(t, dim=axis, keepdim=keepdims)
    return op


def _dims(shape):
    if isinstance(shape, _torch.Tensor):
        return tuple(_b.int(v) for v in shape.reshape(-1))
    return tuple(_b.int(v) for v in shape)


def _reshape(tensor, shape):
    if _is_str(tensor):
        return _np.reshape(_s(tensor), _dims(shape))
    return _t(tensor).reshape(_dims(shape))


def _squeeze(input, axis=None):  # noqa: A002
    t = _t(input)
    if axis is None:
        return t.squeeze()
    assert t.shape[axis] == 1, f"Can not squeeze dim[{axis}], expected a dimension of 1, got {t.shape[axis]}"
    return t.squeeze(axis)


def _expand_dims(input, axis):  # noqa: A002
    return _t(input).unsqueeze(axis)


def _concat(values, axis):
    if _b.any(_is_str(v) for v in values):
        return _np.concatenate([_s(v) for v in values], axis=axis)
    ts = [_t(v) for v in values]
    dt = ts[0].dtype
    for t in ts[1:]:
        dt = _torch.promote_types(dt, t.dtype)
    return _torch.cat([t.to(dt) for t in ts], dim=axis)


def _fill(dims, value):
    if _is_str(value):
        out = _np.empty(_dims(dims), dtype=object)
        out[...] = value.item() if isinstance(value, _np.ndarray) else value
        return out
    v = _t(value)
    return _torch.full(_dims(dims), v.item(), dtype=v.dtype).as_subclass(Tensor)


def _unstack(value, axis=0):
    if isinstance(value, _Shape):
        return list(value)
    return list(_t(value).unbind(axis))


def _range(start, limit=None, delta=1, dtype=None):
    if limit is None:
        start, limit = 0, start
    vals = [_b.float(v) if isinstance(v, _b.float) else (v.item() if isinstance(v, _torch.Tensor) else v)
            for v in (start, limit, delta)]
    if dtype is None:
        dtype = float32 if _b.any(isinstance(v, _b.float) for v in vals) else int32
    return _torch.arange(vals[0], vals[1], vals[2], dtype=dtype).as_subclass(Tensor)


def _matmul(a, b, transpose_a=False, transpose_b=False):
    a, b = _t(a), _t(b)
    if transpose_a:
        a = a.transpose(-1, -2)
    if transpose_b:
        b = b.transpose(-1, -2)
    return _torch.matmul(a, b)


def _transpose(a, perm=None):
    a = _t(a)
    return a.permute(*perm) if perm is not None else a.permute(*reversed(_b.range(a.dim())))


def _softmax(logits, axis=-1):
    return _torch.softmax(_t(logits), dim=axis)


def _binary(fn):
    def op(x, y):
        x = _like(x, y)
        return fn(x, _like(y, x))
    return op


def _unary(fn):
    def op(x):
        return fn(_t(x))
    return op


def _boolean_mask(tensor, mask):
    m = _t(mask).to(_torch.bool)
    if _is_str(tensor):
        return _s(tensor)[m.numpy()]
    return _t(tensor)[m]


def _cumsum(x, axis=0):
    return _torch.cumsum(_t(x), dim=axis)


def _cond(pred, true_fn=None, false_fn=None):
    return true_fn() if _b.bool(pred) else false_fn()


def _gather(params, indices, axis=0):
    idx = _t(indices).to(_torch.int64)
    if _is_str(params):
        return _np.take(_s(params), idx.numpy(), axis=axis)
    return _torch.index_select(_t(params), axis, idx.reshape(-1)).reshape(
        tuple(_t(params).shape[:axis]) + tuple(idx.shape) + tuple(_t(params).shape[axis + 1:]))


def _gather_nd(params, indices):
    """tf.gather_nd; a ragged index tensor gives a ragged result with the same row partition."""
    p = _t(params)
    if isinstance(indices, RaggedTensor):
        idx = _t(indices.values).to(_torch.int64)
        vals = p[tuple(idx[:, j] for j in _b.range(idx.shape[1]))]
        return RaggedTensor(vals, indices.value_rowids, indices.nrows)
    idx = _t(indices).to(_torch.int64)
    return p[tuple(idx[..., j] for j in _b.range(idx.shape[-1]))]


def _tensor_scatter_nd_update(tensor, indices, updates):
    idx = _t(indices).to(_torch.int64)
    assert idx.dim() == 2 and idx.shape[1] == 1, "shim: 1-D scatter only"
    if _is_str(tensor):
        out = _s(tensor).copy()
        out[idx[:, 0].numpy()] = _s(updates)
        return out
    return _torch.index_put(_t(tensor), (idx[:, 0],), _t(updates))


def _sort(values, axis=-1, direction="ASCENDING"):
    return _torch.sort(_t(values), dim=axis, descending=direction != "ASCENDING", stable=True)[0]


def _clip_by_value(t, clip_value_min, clip_value_max):
    t = _t(t)
    return _torch.clamp(t, _like(clip_value_min, t), _like(clip_value_max, t))


def _top_k(input, k=1, sorted=True, name=None):  # noqa: A002
    """tf.math.top_k: descending values; among equal values the lower index comes first."""
    v, i = _torch.sort(_t(input), dim=-1, descending=True, stable=True)
    return v[..., :k], i[..., :k].to(_torch.int32)


def _assert_rank(x, rank, message=None):
    got = _np.ndim(x) if isinstance(x, _np.ndarray) else _t(x).dim()
    if got != rank:
        raise ValueError(message or f"rank {got} != {rank}")


def function(func=None, **_kwargs):
    """tf.function: eager execution of the same Python."""
    if func is None:
        return lambda f: f
    return func


def convert_to_tensor(value, dtype=None, **_kw):
    if _is_str(value):
        return _s(value)
    return _t(value, dtype)


def _scalar_const(fn):
    """float64 mode: a 0-d result is a graph constant (sqrt(d_model), sqrt(depth)) that TensorFlow
    holds as a float32 value - keep that value, like every other float32 constant."""
    def op(x):
        y = fn(_t(x))
        return _round_f32(y).as_subclass(Tensor) if _F64 and y.dim() == 0 and not y.requires_grad else y
    return op


def _pow(x, y):
    x = _like(x, y)
    return _torch.pow(x, _like(y, x))


# --------------------------------------------------------------------------------- namespaces
def _ns(name, **members):
    m = _types.ModuleType(f"tensorflow.{name}")
    m.__dict__.update(members)
    _sys.modules[f"tensorflow.{name}"] = m
    return m


math = _ns(
    "math",
    sqrt=_scalar_const(_torch.sqrt), rsqrt=_scalar_const(_torch.rsqrt), log=_unary(_torch.log), exp=_unary(_torch.exp),
    equal=_equal, logical_not=_unary(_torch.logical_not), logical_and=_binary(_torch.logical_and),
    top_k=_top_k, pow=_pow, minimum=_binary(_torch.minimum), maximum=_binary(_torch.maximum),
    divide=_binary(_torch.true_divide), multiply=_binary(_torch.mul), reduce_sum=_reduce(_torch.sum),
)
nn = _ns("nn", softmax=_softmax, relu=_unary(_torch.relu), sigmoid=_unary(_torch.sigmoid))
debugging = _ns("debugging", assert_rank=_assert_rank)


def _strip(x):
    return _s([v.strip() for v in _s(x).reshape(-1)]).reshape(_np.shape(x))


strings = _ns("strings", strip=_strip)


class _GFile:
    def __init__(self, name, mode="r"):
        self._f = open(name, mode)

    def __enter__(self):
        return self._f

    def __exit__(self, *a):
        self._f.close()


class _Feature:
    def __init__(self, *a, **k):
        self.args, self.kwargs = a, k


io = _ns("io", gfile=_ns("io.gfile", isdir=_os.path.isdir, GFile=_GFile),
         FixedLenFeature=_Feature, VarLenFeature=_Feature)
sparse = _ns("sparse", to_dense=lambda x: x)


# ------------------------------------------------------------------------------- tf.lookup
class _KeyValueTensorInitializer:
    def __init__(self, keys, values, **_kw):
        self.keys = [str(k) for k in _s(keys).reshape(-1)]
        self.values = [_b.int(v) for v in _t(values).reshape(-1)]


class _StaticVocabularyTable:
    """Known keys -> their values; any other key -> len(keys) + hash % num_oov_buckets."""

    def __init__(self, initializer, num_oov_buckets, **_kw):
        self._map = dict(zip(initializer.keys, initializer.values))
        self._n = len(initializer.keys)
        self._oov = num_oov_buckets
        assert num_oov_buckets == 1, "shim: one OOV bucket (any hash lands in it)"

    def size(self):
        return _t(_np.int64(self._n + self._oov))

    def lookup(self, keys):
        a = _s(keys)
        ids = _np.array([self._map.get(k, self._n) for k in a.reshape(-1)], dtype=_np.int64)
        return _torch.from_numpy(ids.reshape(a.shape)).as_subclass(Tensor)


lookup = _ns("lookup", KeyValueTensorInitializer=_KeyValueTensorInitializer,
             StaticVocabularyTable=_StaticVocabularyTable)


# ------------------------------------------------------------------------------- tf.random
class _Random:
    """tf.random.shuffle's stream cannot be matched; the generating script installs the
    permutation it wants (`set_shuffle`) - by default a seeded NumPy permutation."""

    def __init__(self):
        self._rng = _np.random.default_rng(0)
        self._shuffle = None

    def set_shuffle(self, fn):
        self._shuffle = fn

    def shuffle(self, value, seed=None, name=None):
        v = _t(value)
        if self._shuffle is not None:
            perm = _torch.as_tensor(self._shuffle(v.shape[0]), dtype=_torch.int64)
        else:
            perm = _torch.from_numpy(self._rng.permutation(v.shape[0]))
        return v[perm]

    def set_seed(self, seed):
        self._rng = _np.random.default_rng(seed)

    def uniform_like(self, shape):
        return self._rng.random(shape)


random = _Random()


# --------------------------------------------------------------------------------- tf.data
class _Dataset:
    """Just enough of tf.data for examples/BERT4Rec/source/input_pipeline.py:create_cloze_dataset
    with a generator source: lazy, strictly sequential (no prefetch, no parallel map); `shuffle`
    keeps the order (its stream cannot be matched - the generating script feeds the sessions in
    the order it wants them batched)."""

    def __init__(self, make_iter):
        self._make_iter = make_iter

    def __iter__(self):
        return self._make_iter()

    @staticmethod
    def from_generator(generator, output_types=None, output_shapes=None, **_kw):
        def conv(v, dt):
            return _s(v) if _is_str_dtype(dt) else _t(v, dt)

        def it():
            for ex in generator():
                yield {k: conv(ex[k], output_types[k]) for k in output_types}
        return _Dataset(it)

    def shuffle(self, buffer_size, reshuffle_each_iteration=None, seed=None):
        return self

    def repeat(self, count=None):
        def it():
            n = 0
            while count is None or n < count:
                yield from self._make_iter()
                n += 1
        return _Dataset(it)

    def map(self, map_func, num_parallel_calls=None):
        return _Dataset(lambda: (map_func(ex) for ex in self._make_iter()))

    def prefetch(self, buffer_size):
        return self

    def take(self, count):
        def it():
            for i, ex in enumerate(self._make_iter()):
                if i >= count:
                    return
                yield ex
        return _Dataset(it)

    def padded_batch(self, batch_size, padded_shapes=None, padding_values=None, drop_remainder=False):
        def pad(key, items):
            pv = padding_values[key]
            if len(padded_shapes[key]) == 0:
                return _s([str(v.item()) if isinstance(v, _np.ndarray) else v for v in items]) \
                    if _is_str(items[0]) else _torch.stack([_t(v) for v in items]).as_subclass(Tensor)
            width = _b.max(_b.int(_np.shape(v)[0]) if isinstance(v, _np.ndarray) else _b.int(v.shape[0])
                           for v in items)
            if _is_str(items[0]) or (isinstance(items[0], _np.ndarray) and items[0].dtype == object):
                out = _np.empty((len(items), width), dtype=object)
                out[...] = pv
                for r, v in enumerate(items):
                    out[r, :len(v)] = v
                return out
            first = _t(items[0])
            out = _torch.full((len(items), width), _b.float(pv) if first.dtype.is_floating_point else pv,
                              dtype=first.dtype)
            for r, v in enumerate(items):
                out[r, :_t(v).shape[0]] = _t(v)
            return out.as_subclass(Tensor)

        def it():
            buf = []
            for ex in self._make_iter():
                buf.append(ex)
                if len(buf) == batch_size:
                    yield {k: pad(k, [e[k] for e in buf]) for k in padded_shapes}
                    buf = []
            if buf and not drop_remainder:
                yield {k: pad(k, [e[k] for e in buf]) for k in padded_shapes}
        return _Dataset(it)


data = _ns("data", Dataset=_Dataset, experimental=_ns("data.experimental", AUTOTUNE=-1))


# =================================================================================== tf.keras
class _Parameter:
    """A trainable layer weight (float32 / float64 leaf with requires_grad)."""

    def __init__(self, array, name):
        self.name = name
        self.tensor = _torch.tensor(_np.asarray(array, dtype=_np.float32)).to(float32).requires_grad_(True)

    def read(self):
        return self.tensor.as_subclass(Tensor)

    def assign(self, array):
        with _torch.no_grad():
            self.tensor.copy_(_torch.as_tensor(_np.asarray(array, dtype=_np.float32)).to(float32))


_INIT_RNG = _np.random.default_rng(1234)


def set_initializer_seed(seed):
    global _INIT_RNG
    _INIT_RNG = _np.random.default_rng(seed)


def _glorot_uniform(fan_in, fan_out):
    """Keras' default Dense kernel initializer: U(-l, l), l = sqrt(6 / (fan_in + fan_out))."""
    lim = _math.sqrt(6.0 / (fan_in + fan_out))
    return _INIT_RNG.uniform(-lim, lim, size=(fan_in, fan_out)).astype(_np.float32)


class Layer:
    """tf.keras.layers.Layer: __call__ -> call; weights live in `_params` (name -> _Parameter).
    The reference's heads call `super().__init__(kwargs)` with the dict as a positional argument
    (head.py:8, :33, :54): accepted and ignored, as Keras takes it for `trainable`."""

    def __init__(self, *args, **kwargs):
        self._params = {}
        self.name = kwargs.get("name")

    def __call__(self, *args, **kwargs):
        return self.call(*args, **kwargs)

    def add_param(self, name, array):
        self._params[name] = _Parameter(array, name)
        return self._params[name]

    def get_config(self):
        return {"name": self.name}


def tracked_variables(obj, prefix=""):
    """{object-graph path: _Parameter} in the naming of a TF2 object-based checkpoint: attribute
    names from the root, list elements by index, dict entries by key, Sequential members as
    `layer_with_weights-<i>`."""
    out = {}
    seen = set()

    def visit(o, path):
        if id(o) in seen:
            return
        if isinstance(o, Layer):
            seen.add(id(o))
            for n, p in o._params.items():
                out[f"{path}/{n}" if path else n] = p
            if isinstance(o, Sequential):
                for i, l in enumerate(o.layers):
                    visit(l, f"{path}/layer_with_weights-{i}")
                return
            for attr, val in vars(o).items():
                if attr.startswith("_"):
                    continue
                visit(val, f"{path}/{attr}" if path else attr)
        elif isinstance(o, (list, tuple)):
            for i, v in enumerate(o):
                if isinstance(v, (Layer, list, tuple, dict)):
                    visit(v, f"{path}/{i}")
        elif isinstance(o, dict):
            for k, v in o.items():
                if isinstance(v, (Layer, list, tuple, dict)):
                    visit(v, f"{path}/{k}")

    visit(obj, prefix)
    return out


_ACTIVATIONS = {
    None: lambda x: x, "linear": lambda x: x,
    "relu": _torch.relu, "sigmoid": _torch.sigmoid,
    "softmax": lambda x: _torch.softmax(x, dim=-1),
}


class Dense(Layer):
    """outputs = activation(inputs @ kernel + bias); kernel (in, units) glorot-uniform, bias zeros."""

    def __init__(self, units, activation=None, use_bias=True, **kwargs):
        super().__init__(**kwargs)
        self.units = _b.int(units)
        self._activation = activation if callable(activation) else _ACTIVATIONS[activation]
        self._use_bias = use_bias

    def call(self, inputs):
        x = _t(inputs)
        if "kernel" not in self._params:
            self.add_param("kernel", _glorot_uniform(x.shape[-1], self.units))
            if self._use_bias:
                self.add_param("bias", _np.zeros(self.units, _np.float32))
        y = _torch.matmul(x, self._params["kernel"].read())
        if self._use_bias:
            y = y + self._params["bias"].read()
        return self._activation(y)


class Embedding(Layer):
    """Row lookup; `embeddings` (input_dim, output_dim) ~ U(-0.05, 0.05)."""

    def __init__(self, input_dim, output_dim, **kwargs):
        super().__init__(**kwargs)
        self.input_dim, self.output_dim = _b.int(input_dim), _b.int(output_dim)
        self.add_param("embeddings", _INIT_RNG.uniform(-0.05, 0.05, size=(self.input_dim, self.output_dim)))

    def call(self, inputs):
        ids = _t(inputs).to(_torch.int64)
        return self._params["embeddings"].read()[ids]


class LayerNormalization(Layer):
    """Last-axis normalisation.  epsilon 1e-6 < 1.001e-5 rules out Keras' fused kernel, so TF 2.3
    runs nn.moments + nn.batch_normalization: (x - mean) * rsqrt(var + eps) * gamma + beta with
    the biased variance."""

    def __init__(self, axis=-1, epsilon=1e-3, **kwargs):
        super().__init__(**kwargs)
        assert axis == -1
        self.epsilon = epsilon

    def call(self, inputs):
        x = _t(inputs)
        if "gamma" not in self._params:
            self.add_param("gamma", _np.ones(x.shape[-1], _np.float32))
            self.add_param("beta", _np.zeros(x.shape[-1], _np.float32))
        mean = x.mean(dim=-1, keepdim=True)
        var = ((x - mean) ** 2).mean(dim=-1, keepdim=True)
        inv = _torch.rsqrt(var + self.epsilon) * self._params["gamma"].read()
        return x * inv + (self._params["beta"].read() - mean * inv)


DROPOUT_LOG = []          # (layer, scaled mask as NumPy) per training-mode call, in call order


class Dropout(Layer):
    """training: x * keep / (1 - rate), keep = uniform >= rate (tf.nn.dropout); else identity."""

    def __init__(self, rate, **kwargs):
        super().__init__(**kwargs)
        self.rate = _b.float(rate)

    def call(self, inputs, training=None):
        x = _t(inputs)
        if not training or self.rate == 0.0:
            return x
        keep = random.uniform_like(tuple(x.shape)) >= self.rate
        mask = _torch.from_numpy(keep.astype(_np.float32) / _np.float32(1.0 - self.rate)).to(x.dtype)
        DROPOUT_LOG.append((self, mask.numpy().astype(_np.float32)))
        return x * mask


class Sequential(Layer):
    def __init__(self, layers=None, **kwargs):
        super().__init__(**kwargs)
        self.layers = list(layers or [])

    def call(self, inputs, training=None, mask=None):
        x = inputs
        for l in self.layers:
            x = l(x)
        return x


class Model(Layer):
    pass


class _Loss:
    """tf.keras.losses.Loss with Reduction.NONE: __call__ returns call()'s value unchanged."""

    def __init__(self, reduction="auto", name=None):
        self.reduction, self.name = reduction, name

    def __call__(self, y_true, y_pred, sample_weight=None):
        assert sample_weight is None
        return self.call(_t(y_true), _t(y_pred))


class _Metric(Layer):
    def __init__(self, name=None, dtype=None, **kwargs):
        super().__init__(name=name)

    def add_weight(self, name=None, shape=(), initializer="zeros", dtype=None, **_kw):
        assert initializer == "zeros"
        return Variable(_torch.zeros(shape, dtype=float32), name=name)

    def __call__(self, *args, **kwargs):
        self.update_state(*args, **kwargs)
        return self.result()


_EPS = 1e-7   # tf.keras.backend.epsilon()


def _k_sparse_categorical_crossentropy(target, output, from_logits=False, axis=-1):
    """tf.keras.backend.sparse_categorical_crossentropy on an eager tensor (TF 2.3): unless
    from_logits, output = log(clip(output, eps, 1 - eps)); target flattened to int64 when its rank
    is not rank(output) - 1; then sparse_softmax_cross_entropy_with_logits - i.e. a SECOND
    log-softmax over the clipped log-probabilities (a no-op while they sum to one)."""
    out = _t(output)
    tgt = _t(target).to(_torch.int64)
    assert axis in (-1, out.dim() - 1)
    if not from_logits:
        out = _torch.log(_torch.clamp(out, _EPS, 1.0 - _EPS))
    out_rank = out.dim()
    out_shape = tuple(out.shape)
    update_shape = tgt.dim() != out_rank - 1
    if update_shape:
        tgt = tgt.reshape(-1)
        out = out.reshape(-1, out.shape[-1])
    logp = _torch.log_softmax(out, dim=-1)
    res = -_torch.gather(logp, -1, tgt.unsqueeze(-1)).squeeze(-1)
    if update_shape and out_rank >= 3:
        res = res.reshape(out_shape[:-1])
    return res


def _k_binary_crossentropy(target, output, from_logits=False):
    """tf.keras.backend.binary_crossentropy on an eager tensor (TF 2.3): clip to [eps, 1 - eps],
    -(t log(o + eps) + (1 - t) log(1 - o + eps))."""
    assert not from_logits
    t = _t(target)
    o = _torch.clamp(_t(output), _EPS, 1.0 - _EPS)
    bce = t * _torch.log(o + _EPS)
    bce = bce + (1 - t) * _torch.log(1 - o + _EPS)
    return -bce


def _losses_binary_crossentropy(y_true, y_pred, from_logits=False, label_smoothing=0):
    return _k_binary_crossentropy(_t(y_true, _t(y_pred).dtype), y_pred, from_logits).mean(dim=-1)


class _LearningRateSchedule:
    pass


class _Callback:
    def __init__(self, *a, **k):
        self.model = None


class _BinaryCrossentropy(_Loss):
    def call(self, y_true, y_pred):
        return _losses_binary_crossentropy(y_true, y_pred).mean()


keras = _ns(
    "keras",
    layers=_ns("keras.layers", Layer=Layer, Dense=Dense, Embedding=Embedding, Dropout=Dropout,
               LayerNormalization=LayerNormalization),
    models=_ns("keras.models", Model=Model),
    Model=Model,
    Sequential=Sequential,
    activations=_ns("keras.activations", softmax=lambda x, axis=-1: _torch.softmax(_t(x), dim=axis),
                    relu=_torch.relu, sigmoid=_torch.sigmoid),
    losses=_ns("keras.losses", Loss=_Loss, Reduction=_types.SimpleNamespace(NONE="none", AUTO="auto"),
               binary_crossentropy=_losses_binary_crossentropy, BinaryCrossentropy=_BinaryCrossentropy),
    metrics=_ns("keras.metrics", Metric=_Metric),
    backend=_ns("keras.backend", sparse_categorical_crossentropy=_k_sparse_categorical_crossentropy,
                binary_crossentropy=_k_binary_crossentropy, epsilon=lambda: _EPS,
                eval=lambda x: _t(x).numpy()),
    optimizers=_ns("keras.optimizers", schedules=_ns(
        "keras.optimizers.schedules", LearningRateSchedule=_LearningRateSchedule)),
    callbacks=_ns("keras.callbacks", Callback=_Callback, TensorBoard=_Callback, ModelCheckpoint=_Callback),
)

# names that shadow builtins come last
cast = _cast
shape = _shape
size = _size
equal = _equal
where = _where
reduce_sum = _reduce(_torch.sum)
reduce_mean = _reduce(_torch.mean)
reduce_max = _reduce(_torch.amax)
reshape = _reshape
squeeze = _squeeze
expand_dims = _expand_dims
concat = _concat
fill = _fill
unstack = _unstack
matmul = _matmul
transpose = _transpose
multiply = _binary(_torch.mul)
divide = _binary(_torch.true_divide)
logical_not = _unary(_torch.logical_not)
logical_and = _binary(_torch.logical_and)
boolean_mask = _boolean_mask
cumsum = _cumsum
cond = _cond
gather = _gather
gather_nd = _gather_nd
tensor_scatter_nd_update = _tensor_scatter_nd_update
sort = _sort
clip_by_value = _clip_by_value
round = _unary(_torch.round)      # noqa: A001  (half to even, as tf.round)
range = _range                    # noqa: A001
bool = _bool                      # noqa: A001
