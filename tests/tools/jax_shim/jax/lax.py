"""jax.lax control flow, eager."""
import numpy as _np

from ._core import Arr, asarr
from . import tree_util as _tu


def _truth(p):
    return bool(_np.asarray(p).reshape(-1)[0]) if _np.asarray(p).size == 1 else bool(p)


def select(pred, on_true, on_false):
    p = _np.asarray(pred)
    a, b = _np.asarray(on_true), _np.asarray(on_false)
    if p.ndim == 0:
        r = a if _truth(p) else b
        dt = _np.result_type(a, b) if a.dtype != b.dtype else a.dtype
        if isinstance(on_true, (int, float, bool)) and not isinstance(on_false, (int, float, bool)):
            dt = b.dtype
        if isinstance(on_false, (int, float, bool)) and not isinstance(on_true, (int, float, bool)):
            dt = a.dtype
        return asarr(_np.array(r, copy=True), dt)
    return asarr(_np.where(p, a, b))


def cond(pred, true_fun, false_fun, *operands):
    return true_fun(*operands) if _truth(pred) else false_fun(*operands)


def switch(index, branches, *operands):
    i = int(_np.clip(int(_np.asarray(index)), 0, len(branches) - 1))
    return branches[i](*operands)


def while_loop(cond_fun, body_fun, init_val):
    val = init_val
    while _truth(cond_fun(val)):
        val = body_fun(val)
    return val


def fori_loop(lower, upper, body_fun, init_val):
    val = init_val
    for i in range(int(lower), int(upper)):
        val = body_fun(asarr(_np.array(i, dtype=_np.int32)), val)
    return val


def scan(f, init, xs, length=None):
    carry = init
    ys = []
    n = length if xs is None else _tu.tree_leaves(xs)[0].shape[0]
    for i in range(n):
        x = None if xs is None else _tu.tree_map(lambda a: a[i], xs)
        carry, y = f(carry, x)
        ys.append(y)
    return carry, _tu.tree_stack(ys)


def map(f, xs):  # noqa: A001
    n = _tu.tree_leaves(xs)[0].shape[0]
    return _tu.tree_stack([f(_tu.tree_map(lambda a: a[i], xs)) for i in range(n)])


def convert_element_type(x, dtype):
    return asarr(x, dtype)
