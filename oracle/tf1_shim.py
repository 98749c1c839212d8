"""TEST INFRASTRUCTURE ONLY -- eager numpy stand-in for the TensorFlow-1 symbols the reference's hot path calls.

Purpose: let the reference's own graph-building Python (``slam_recognition/filters/*.py``, ``util/apply_filter.py``,
``util/regulator/gaussian_regulator_tensor.py``, ``util/selection/*.py``, ``util/color/get_value.py``) run UNMODIFIED in
the build container, where TensorFlow cannot be installed, so that ``tests/golden/make_golden.py`` can record what that
code computes. It is never imported by the product package.

PARITY UNPINNED: the op semantics below are restated from TF-1 documentation; no reference test
pins any result at this boundary and real TensorFlow is unavailable to cross-check.

Restated semantics (NHWC, float32 tensors):

* ``tf.nn.conv2d`` : cross-correlation, filter ``[kh, kw, Cin, Cout]``; ``SAME``: ``out = ceil(n / stride)``,
  ``pad_total = max((out - 1) * stride + k - n, 0)``, ``pad_before = pad_total // 2``, zero fill. Accumulated in float64
  and rounded once to float32 (the correctly-rounded value any float32 implementation approximates).
* ``tf.nn.max_pool`` ``SAME``: same geometry, padding ignored (window clipped to the image).
* ``tf.image.resize_images(NEAREST_NEIGHBOR)`` (align_corners=False): ``src = min(floor(dst * float32(in / out)), in - 1)``.
* ``tf.where(cond)`` -> int64 ``[K, rank]`` in row-major order; ``tf.where(cond, a, b)`` elementwise select.
* ``tf.maximum`` / ``tf.minimum`` / ``tf.clip_by_value`` / ``max_pool``: NaN-PROPAGATING (TF-1's behaviour is
  Eigen-packet dependent, i.e. undefined; this is the documented choice).
* ``tf.pow`` : float32 ``powf`` (computed in float64, rounded to float32).
"""
import sys
import types

import numpy as np

float32 = np.float32
int32 = np.int32
int64 = np.int64


class Shape(tuple):
    """TensorShape stand-in: ints, sliceable, ``len``-able, and ``shape - tensor`` yields an int array."""

    def __new__(cls, dims):
        return super().__new__(cls, [int(d) for d in dims])

    def __getitem__(self, key):
        got = super().__getitem__(key)
        return Shape(got) if isinstance(key, slice) else got

    def __sub__(self, other):
        return Tensor(np.asarray(self, dtype=np.int32) - _arr(other).astype(np.int32))

    def as_list(self):
        return list(self)

    def concatenate(self, other):
        return Shape(list(self) + [int(d) for d in other])


TensorShape = Shape


class Tensor:
    __array_priority__ = 1000

    def __init__(self, value):
        self._v = np.asarray(value)

    def numpy(self):
        return self._v

    def set_shape(self, shape):   # static-shape hint in TF: nothing to do for an eager array
        assert tuple(int(d) for d in shape) == self._v.shape, (tuple(shape), self._v.shape)

    def value(self):
        return Tensor(self._v)

    @property
    def shape(self):
        return Shape(self._v.shape)

    def get_shape(self):
        return self.shape

    @property
    def dtype(self):
        return self._v.dtype

    def __getitem__(self, key):
        return Tensor(self._v[key])

    def __len__(self):
        return len(self._v)

    def _bin(self, other, fn, swap=False):
        a, b = self._v, _arr(other)
        if b.dtype != a.dtype and (isinstance(other, (int, float, list, tuple)) or b.dtype.kind != a.dtype.kind):
            b = b.astype(a.dtype)
        return Tensor(fn(b, a) if swap else fn(a, b))

    def __add__(self, o): return self._bin(o, np.add)
    def __radd__(self, o): return self._bin(o, np.add, True)
    def __sub__(self, o): return self._bin(o, np.subtract)
    def __rsub__(self, o): return self._bin(o, np.subtract, True)
    def __mul__(self, o): return self._bin(o, np.multiply)
    def __rmul__(self, o): return self._bin(o, np.multiply, True)
    def __truediv__(self, o): return self._bin(o, _divide)
    def __rtruediv__(self, o): return self._bin(o, _divide, True)
    def __neg__(self): return Tensor(-self._v)
    def __pow__(self, o): return pow(self, o)


def _divide(a, b):
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.divide(a, b)


def _arr(x):
    if isinstance(x, Tensor):
        return x._v
    return np.asarray(x)


def _f32(x):
    return _arr(x).astype(np.float32)


def constant(value, dtype=None, shape=None):
    a = np.asarray(value)
    if dtype is not None:
        a = a.astype(dtype)
    if shape is not None:
        a = a.reshape(tuple(shape))
    return Tensor(a)


def cast(x, dtype):
    return Tensor(_arr(x).astype(dtype))


def shape(x):
    return Tensor(np.asarray(_arr(x).shape, dtype=np.int32))


def ones(shp, dtype=np.float32):
    return Tensor(np.ones(tuple(int(d) for d in _arr(shp).reshape(-1)), dtype=dtype))


def ones_like(x):
    return Tensor(np.ones_like(_arr(x)))


def zeros_like(x):
    return Tensor(np.zeros_like(_arr(x)))


def pad(x, paddings):
    return Tensor(np.pad(_arr(x), [tuple(int(v) for v in p) for p in _arr(paddings)], mode="constant"))


def reduce_sum(x, axis=None, keepdims=False):
    a = _arr(x)
    if a.dtype == np.float32:
        # TF reduces along the axis in float32; for the 3-channel sums on this path the order is ((c0 + c1) + c2).
        ax = axis if axis is not None else tuple(range(a.ndim))
        if isinstance(ax, int) and a.shape[ax] <= 8:
            parts = np.moveaxis(a, ax, 0)
            acc = parts[0].copy()
            for p in parts[1:]:
                acc = acc + p
            return Tensor(np.expand_dims(acc, ax) if keepdims else acc)
    return Tensor(np.sum(a, axis=axis, keepdims=keepdims))


def _nanmax2(a, b):
    return np.maximum(a, b)   # numpy propagates NaN


def maximum(x, y):
    a = _arr(x)
    return Tensor(np.maximum(a, _arr(y).astype(a.dtype)))


def minimum(x, y):
    a = _arr(x)
    return Tensor(np.minimum(a, _arr(y).astype(a.dtype)))


def clip_by_value(x, lo, hi):
    a = _arr(x)
    return Tensor(np.minimum(np.maximum(a, a.dtype.type(lo)), a.dtype.type(hi)))


def pow(x, y):  # noqa: A001  (mirrors tf.pow)
    a = _arr(x)
    e = _arr(y).astype(a.dtype)
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        if a.dtype == np.float32:
            return Tensor(np.power(a.astype(np.float64), e.astype(np.float64)).astype(np.float32))
        return Tensor(np.power(a, e))


def greater_equal(x, y):
    with np.errstate(invalid="ignore"):
        return Tensor(_arr(x) >= _arr(y))


def equal(x, y):
    return Tensor(_arr(x) == _arr(y))


def where(cond, x=None, y=None):
    c = _arr(cond)
    if x is None:
        return Tensor(np.argwhere(c).astype(np.int64))
    return Tensor(np.where(c, _arr(x), _arr(y)))


def abs(x):  # noqa: A001
    return Tensor(np.abs(_arr(x)))


class Variable(Tensor):
    """``tf.Variable``: a tensor with state; ``assign`` stores and returns the new value."""

    def __init__(self, initial_value, dtype=None, **kw):
        a = _arr(initial_value)
        super().__init__(a.astype(dtype) if dtype is not None else a.copy())

    def assign(self, value):
        self._v = _arr(value).astype(self._v.dtype)
        return Tensor(self._v)


def convolution(input=None, filter=None, strides=None, padding="SAME", **kw):  # noqa: A002
    """``tf.nn.convolution`` for NHWC rank-4 input: ``strides`` lists the SPATIAL strides only."""
    sp = [1, 1] if strides is None else [int(v) for v in strides]
    return conv2d(input=input, filter=filter, strides=(1, sp[0], sp[1], 1), padding=padding)


def less_equal(x, y):
    with np.errstate(invalid="ignore"):
        return Tensor(_arr(x) <= _arr(y))


# ---- tensorflow.python.ops.array_ops / framework.ops symbols used by util/color/to_channels.py --------------------------
class _NameScope:
    def __init__(self, name, default_name=None, values=None):
        self.name = name or default_name

    def __enter__(self):
        return self.name

    def __exit__(self, *exc):
        return False


def convert_to_tensor(x, name=None, **kw):
    return x if isinstance(x, Tensor) else Tensor(np.asarray(x))


def rank(x):
    return Tensor(np.asarray(_arr(x).ndim, dtype=np.int32))


def expand_dims(x, axis):
    return Tensor(np.expand_dims(_arr(x), int(axis)))


def concat(values, axis=0, **kw):
    return Tensor(np.concatenate([np.atleast_1d(_arr(v)) for v in values], axis=int(axis)))


def tile(x, multiples, name=None):
    return Tensor(np.tile(_arr(x), tuple(int(m) for m in _arr(multiples).reshape(-1))))


def _same_geometry(n, k, s):
    out = -(-n // s)
    pad_total = max((out - 1) * s + k - n, 0)
    return out, pad_total // 2


def conv2d(input=None, filter=None, strides=(1, 1, 1, 1), padding="SAME", **kw):  # noqa: A002
    x = _f32(input)
    w = _f32(filter)
    assert padding == "SAME" and x.ndim == 4 and w.ndim == 4 and x.shape[3] == w.shape[2]
    n, h, wd, _ = x.shape
    kh, kw_, _, cout = w.shape
    sy, sx = int(strides[1]), int(strides[2])
    oh, pt = _same_geometry(h, kh, sy)
    ow, pl = _same_geometry(wd, kw_, sx)
    need_h = (oh - 1) * sy + kh
    need_w = (ow - 1) * sx + kw_
    xp = np.zeros((n, max(need_h, pt + h), max(need_w, pl + wd), x.shape[3]), dtype=np.float64)
    xp[:, pt:pt + h, pl:pl + wd, :] = x
    acc = np.zeros((n, oh, ow, cout), dtype=np.float64)
    w64 = w.astype(np.float64)
    with np.errstate(invalid="ignore", over="ignore"):
        for ky in range(kh):
            for kx in range(kw_):
                patch = xp[:, ky:ky + (oh - 1) * sy + 1:sy, kx:kx + (ow - 1) * sx + 1:sx, :]
                acc += patch @ w64[ky, kx]
    return Tensor(acc.astype(np.float32))


def max_pool(value, ksize, strides=None, padding="SAME", **kw):
    x = _arr(value)
    assert padding == "SAME" and x.ndim == 4
    n, h, wd, c = x.shape
    kh, kw_ = int(ksize[1]), int(ksize[2])
    sy, sx = int(strides[1]), int(strides[2])
    oh, pt = _same_geometry(h, kh, sy)
    ow, pl = _same_geometry(wd, kw_, sx)
    out = np.empty((n, oh, ow, c), dtype=x.dtype)
    for i in range(oh):
        y0, y1 = max(i * sy - pt, 0), min(i * sy - pt + kh, h)
        for j in range(ow):
            x0, x1 = max(j * sx - pl, 0), min(j * sx - pl + kw_, wd)
            out[:, i, j, :] = np.max(x[:, y0:y1, x0:x1, :], axis=(1, 2))   # np.max propagates NaN
    return Tensor(out)


class _ResizeMethod:
    NEAREST_NEIGHBOR = 1


def _resize_nearest(images, size, **kw):
    x = _arr(images)
    oh, ow = (int(s) for s in _arr(size).reshape(-1))
    h, wd = x.shape[1:3]
    ys = np.minimum(np.floor(np.arange(oh, dtype=np.float32) * np.float32(h / oh)).astype(np.int64), h - 1)
    xs = np.minimum(np.floor(np.arange(ow, dtype=np.float32) * np.float32(wd / ow)).astype(np.int64), wd - 1)
    return Tensor(x[:, ys][:, :, xs])


def resize_images(images, size, method=None, **kw):
    assert method == _ResizeMethod.NEAREST_NEIGHBOR
    return _resize_nearest(images, size)


def grayscale_to_rgb(x):
    return Tensor(np.repeat(_arr(x), 3, axis=-1))


def install():
    """Register stub ``tensorflow*`` modules in ``sys.modules`` (idempotent)."""
    if "tensorflow" in sys.modules and getattr(sys.modules["tensorflow"], "__silent_shim__", False):
        return sys.modules["tensorflow"]
    tf = types.ModuleType("tensorflow")
    tf.__silent_shim__ = True
    for name in ("Tensor", "TensorShape", "float32", "int32", "int64", "constant", "cast", "shape", "ones", "ones_like",
                 "zeros_like", "pad", "reduce_sum", "maximum", "minimum", "clip_by_value", "pow", "greater_equal",
                 "equal", "where", "abs", "Variable"):
        setattr(tf, name, globals()[name])
    tf.nn = types.SimpleNamespace(conv2d=conv2d, max_pool=max_pool, convolution=convolution)
    tf.math = types.SimpleNamespace(less_equal=less_equal)
    tf.image = types.SimpleNamespace(resize_images=resize_images, resize_nearest_neighbor=_resize_nearest,
                                     ResizeMethod=_ResizeMethod, grayscale_to_rgb=grayscale_to_rgb)
    mods = {"tensorflow": tf}
    for sub in ("python", "python.ops", "python.ops.math_ops", "python.ops.array_ops", "python.framework",
                "python.framework.ops", "python.framework.dtypes"):
        m = types.ModuleType("tensorflow." + sub)
        mods["tensorflow." + sub] = m
    mods["tensorflow.python.framework.dtypes"].int32 = np.int32
    for name in ("expand_dims", "rank", "ones", "concat", "tile"):
        setattr(mods["tensorflow.python.ops.array_ops"], name, globals()[name])
    mods["tensorflow.python.framework.ops"].name_scope = _NameScope
    mods["tensorflow.python.framework.ops"].convert_to_tensor = convert_to_tensor
    for full, m in mods.items():
        sys.modules[full] = m
        if "." in full:
            parent, _, leaf = full.rpartition(".")
            setattr(mods[parent], leaf, m)
    return tf
