"""jax.numpy subset, eager on NumPy (int32 / float32 defaults)."""
import builtins as _bi

import numpy as _np

from ._core import Arr, asarr, canon_dtype

int32 = _np.int32
uint32 = _np.uint32
int8 = _np.int8
uint8 = _np.uint8
float32 = _np.float32
bool_ = _np.bool_
inf = float("inf")
ndarray = Arr


def _w(x):
    return asarr(x)


def array(x, dtype=None, copy=True):
    if isinstance(x, (list, tuple)):
        x = _np.array([_np.asarray(e) for e in x]) if len(x) and not _np.isscalar(x[0]) else _np.array(x)
    return asarr(_np.array(x, copy=True), dtype)


asarray = array


def zeros(shape, dtype=float):
    return asarr(_np.zeros(_shape(shape), canon_dtype(dtype)))


def ones(shape, dtype=float):
    return asarr(_np.ones(_shape(shape), canon_dtype(dtype)))


def full(shape, fill_value, dtype=None):
    a = _np.full(_shape(shape), _np.asarray(fill_value))
    return asarr(a, dtype)


def _shape(shape):
    if isinstance(shape, (tuple, list)):
        return tuple(int(s) for s in shape)
    return int(shape)


def arange(*args, dtype=None):
    return asarr(_np.arange(*[int(a) if float(a).is_integer() else a for a in args]), dtype)


def stack(arrays, axis=0):
    return _w(_np.stack([_np.asarray(a) for a in arrays], axis=axis))


def concatenate(arrays, axis=0):
    return _w(_np.concatenate([_np.asarray(a) for a in arrays], axis=axis))


def hstack(arrays):
    return _w(_np.hstack([_np.atleast_1d(_np.asarray(a)) for a in arrays]))


def divmod(a, b):  # noqa: A001
    q, r = _np.divmod(_np.asarray(a), _np.asarray(b))
    return _w(q), _w(r)


def where(c, a=None, b=None):
    if a is None:
        return tuple(_w(i) for i in _np.where(_np.asarray(c)))
    a_, b_ = _np.asarray(a), _np.asarray(b)
    # python scalars are weakly typed, as in jax
    if isinstance(a, (int, float, bool)) and not isinstance(b, (int, float, bool)):
        a_ = a_.astype(_np.result_type(b_.dtype, _np.float32) if isinstance(a, float) and b_.dtype.kind in "iub" else b_.dtype)
    if isinstance(b, (int, float, bool)) and not isinstance(a, (int, float, bool)):
        b_ = b_.astype(_np.result_type(a_.dtype, _np.float32) if isinstance(b, float) and a_.dtype.kind in "iub" else a_.dtype)
    return _w(_np.where(_np.asarray(c), a_, b_))


def argwhere(a, size=None, fill_value=0):
    r = _np.argwhere(_np.asarray(a))
    if size is not None:
        out = _np.full((size, r.shape[1]), fill_value, dtype=_np.int32)
        n = _bi.min(size, r.shape[0])
        out[:n] = r[:n]
        r = out
    return asarr(r, _np.int32)


def max(a, axis=None):  # noqa: A001
    return _w(_np.max(_np.asarray(a), axis=axis))


def min(a, axis=None):  # noqa: A001
    return _w(_np.min(_np.asarray(a), axis=axis))


def sum(a, axis=None):  # noqa: A001
    a = _np.asarray(a)
    r = _np.sum(a, axis=axis)
    return asarr(r, _np.int32 if a.dtype.kind in "biu" else None)


def mean(a, axis=None):
    return asarr(_np.mean(_np.asarray(a), axis=axis, dtype=_np.float32), _np.float32)


def cumsum(a, axis=None):
    return _w(_np.cumsum(_np.asarray(a), axis=axis, dtype=_np.asarray(a).dtype))


def argmax(a, axis=None):
    return asarr(_np.argmax(_np.asarray(a), axis=axis), _np.int32)


def argmin(a, axis=None):
    return asarr(_np.argmin(_np.asarray(a), axis=axis), _np.int32)


def any(a, axis=None):  # noqa: A001
    return _w(_np.any(_np.asarray(a), axis=axis))


def all(a, axis=None):  # noqa: A001
    return _w(_np.all(_np.asarray(a), axis=axis))


def logical_and(a, b):
    return _w(_np.logical_and(_np.asarray(a), _np.asarray(b)))


def logical_or(a, b):
    return _w(_np.logical_or(_np.asarray(a), _np.asarray(b)))


def logical_not(a):
    return _w(_np.logical_not(_np.asarray(a)))


def array_equal(a, b):
    a, b = _np.asarray(a), _np.asarray(b)
    return _w(_np.array(a.shape == b.shape and bool((a == b).all())))


def flip(a, axis=None):
    return _w(_np.flip(_np.asarray(a), axis=axis))


def meshgrid(*xs, indexing="xy"):
    return [_w(m) for m in _np.meshgrid(*[_np.asarray(x) for x in xs], indexing=indexing)]


def take(a, ind, axis=None, **kw):
    a = _np.asarray(a)
    ind = _np.asarray(ind)
    n = a.shape[axis] if axis is not None else a.size
    ind = _np.clip(_np.where(ind < 0, ind + n, ind), 0, n - 1)
    return _w(_np.take(a, ind, axis=axis))


def searchsorted(a, v, side="left"):
    return asarr(_np.searchsorted(_np.asarray(a), _np.asarray(v), side=side), _np.int32)


def ravel(a):
    return _w(_np.ravel(_np.asarray(a)))


def reshape(a, shape):
    return _w(_np.reshape(_np.asarray(a), shape))


def zeros_like(a, dtype=None):
    return asarr(_np.zeros_like(_np.asarray(a)), dtype)
