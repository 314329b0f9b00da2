"""Array type with JAX indexing semantics on top of numpy.ndarray."""
import numpy as np

_INT = (int, np.integer)


def canon_dtype(dtype):
    """x64 disabled: python int -> int32, float -> float32."""
    if dtype is None:
        return None
    if dtype is int:
        return np.dtype(np.int32)
    if dtype is float:
        return np.dtype(np.float32)
    if dtype is bool:
        return np.dtype(np.bool_)
    d = np.dtype(dtype)
    if d == np.int64:
        return np.dtype(np.int32)
    if d == np.uint64:
        return np.dtype(np.uint32)
    if d == np.float64:
        return np.dtype(np.float32)
    return d


def asarr(x, dtype=None):
    dtype = canon_dtype(dtype)
    a = np.asarray(x)
    if dtype is None:
        dtype = canon_dtype(a.dtype)
    if a.dtype != dtype:
        a = a.astype(dtype)
    return a.view(Arr)


def _is_int_index(i):
    return isinstance(i, _INT) or (isinstance(i, np.ndarray) and i.dtype.kind in "iu")


def _gather_index(shape, idx):
    """negative indices wrap once, then everything is clamped (jnp gather default)."""
    if not isinstance(idx, tuple):
        idx = (idx,)
    out, dim = [], 0
    for i in idx:
        if i is None or i is Ellipsis:
            if i is Ellipsis:
                dim = len(shape) - (len(idx) - len(out) - 1)
            out.append(i)
            continue
        if isinstance(i, (list, tuple)):
            i = np.asarray(i)
        if _is_int_index(i):
            n = shape[dim]
            a = np.asarray(i).astype(np.int64)
            a = np.where(a < 0, a + n, a)
            a = np.clip(a, 0, n - 1)
            out.append(int(a) if a.ndim == 0 else a)
            dim += 1
        elif isinstance(i, np.ndarray) and i.dtype == np.bool_:
            out.append(np.asarray(i))
            dim += i.ndim
        else:
            out.append(i)
            dim += 1
    return tuple(out)


class _AtIndexer:
    def __init__(self, arr):
        self.arr = arr

    def __getitem__(self, idx):
        return _AtRef(self.arr, idx)


class _AtRef:
    def __init__(self, arr, idx):
        self.arr, self.idx = arr, idx

    def _resolve(self):
        """-> (tuple of broadcast int64 index arrays, validity mask) for pure integer indexing, or None."""
        idx = self.idx if isinstance(self.idx, tuple) else (self.idx,)
        idx = tuple(np.asarray(i) if isinstance(i, (list, tuple)) else i for i in idx)
        if not all(_is_int_index(i) for i in idx):
            return None
        shape = self.arr.shape
        arrs = np.broadcast_arrays(*[np.asarray(i).astype(np.int64) for i in idx])
        valid = np.ones(arrs[0].shape, dtype=bool)
        norm = []
        for d, a in enumerate(arrs):
            n = shape[d]
            a = np.where(a < 0, a + n, a)
            valid &= (a >= 0) & (a < n)
            norm.append(np.clip(a, 0, n - 1))
        return tuple(norm), valid

    def _apply(self, values, op):
        out = np.array(self.arr, copy=True)
        res = self._resolve()
        if res is None:  # slices / masks: plain numpy semantics
            if op == "set":
                out[self.idx] = values
            else:
                np.add.at(out, self.idx, values)
            return out.view(Arr)
        norm, valid = res
        nidx = len(norm)
        tail = out.shape[nidx:]
        vals = np.broadcast_to(np.asarray(values).astype(out.dtype), valid.shape + tail)
        if valid.ndim == 0:
            if bool(valid):
                key = tuple(int(a) for a in norm)
                if op == "set":
                    out[key] = vals
                else:
                    out[key] += vals
            return out.view(Arr)
        flat = [a.reshape(-1) for a in norm]
        fv = valid.reshape(-1)
        vflat = vals.reshape((-1,) + tail)
        for j in range(fv.shape[0]):  # in order: a later duplicate wins, out-of-range updates are dropped
            if not fv[j]:
                continue
            key = tuple(int(a[j]) for a in flat)
            if op == "set":
                out[key] = vflat[j]
            else:
                out[key] += vflat[j]
        return out.view(Arr)

    def set(self, values):
        return self._apply(values, "set")

    def add(self, values):
        return self._apply(values, "add")


class Arr(np.ndarray):
    """numpy array with `.at[...]`, clamped integer gathers and int32/float32 results."""

    __array_priority__ = 100

    @property
    def at(self):
        return _AtIndexer(self)

    def __getitem__(self, idx):
        try:
            nidx = _gather_index(self.shape, idx)
        except Exception:
            nidx = idx
        r = np.ndarray.__getitem__(self, nidx)
        if not isinstance(r, np.ndarray):  # full integer index: keep a 0-d array, like jax
            r = np.asarray(r).view(Arr)
        return r

    def __array_finalize__(self, obj):
        pass

    def __array_wrap__(self, out, context=None, return_scalar=False):
        r = np.asarray(out)
        d = canon_dtype(r.dtype)
        if r.dtype != d:
            r = r.astype(d)
        if r.ndim == 0 and return_scalar:
            return r.view(Arr)
        return r.view(Arr)

    def __hash__(self):
        return id(self)

    def __bool__(self):
        return bool(np.asarray(self).reshape(-1)[0]) if self.size == 1 else np.ndarray.__bool__(self)

    def __iter__(self):
        if self.ndim == 0:
            raise TypeError("iteration over a 0-d array")
        for i in range(self.shape[0]):
            yield self[i]

    def __index__(self):
        return int(np.asarray(self))
