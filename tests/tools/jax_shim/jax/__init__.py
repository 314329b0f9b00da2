"""Eager NumPy stand-in for the parts of jax the reference's hot-path files use (see ../README.md)."""
from . import numpy, lax, random, debug, tree_util  # noqa: F401
from ._core import Arr as Array  # noqa: F401
from ._vmap import vmap  # noqa: F401


def jit(fun=None, **kwargs):
    if fun is None:
        return lambda f: f
    return fun


class disable_jit:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def __call__(self, f):
        return f


def block_until_ready(x):
    return x
