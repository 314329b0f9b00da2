"""chex stand-in: type aliases and `dataclass` (a plain dataclass with `replace`)."""
import dataclasses as _dc
from typing import Any

Array = Any
PRNGKey = Any
Numeric = Any


def dataclass(cls=None, **kwargs):
    def wrap(c):
        eq = "__eq__" in c.__dict__
        d = _dc.dataclass(c, eq=not eq)
        d.replace = lambda self, **kw: _dc.replace(self, **kw)
        return d

    return wrap if cls is None else wrap(cls)
