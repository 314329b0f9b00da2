"""Minimal pytrees: tuple, list, dict, NamedTuple, dataclasses; everything else is a leaf."""
import dataclasses

import numpy as _np

from ._core import asarr


def _is_namedtuple(x):
    return isinstance(x, tuple) and hasattr(x, "_fields")


def tree_map(f, tree, *rest):
    if _is_namedtuple(tree):
        return type(tree)(*[tree_map(f, t, *[r[i] for r in rest]) for i, t in enumerate(tree)])
    if isinstance(tree, (tuple, list)):
        return type(tree)(tree_map(f, t, *[r[i] for r in rest]) for i, t in enumerate(tree))
    if isinstance(tree, dict):
        return {k: tree_map(f, v, *[r[k] for r in rest]) for k, v in tree.items()}
    if dataclasses.is_dataclass(tree) and not isinstance(tree, type):
        return type(tree)(**{fl.name: tree_map(f, getattr(tree, fl.name), *[getattr(r, fl.name) for r in rest]) for fl in dataclasses.fields(tree)})
    if tree is None:
        return None
    return f(tree, *rest)


def tree_leaves(tree):
    out = []
    tree_map(lambda x: out.append(x), tree)
    return out


def tree_stack(trees):
    if not trees:
        return None
    return tree_map(lambda *xs: asarr(_np.stack([_np.asarray(x) for x in xs])), trees[0], *trees[1:])
