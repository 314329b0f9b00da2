"""jax.vmap as a python loop over axis 0 (the only axis the reference maps)."""
import numpy as _np

from . import tree_util as _tu


def vmap(fun, in_axes=0, out_axes=0, **_):
    assert out_axes == 0

    def mapped(*args, **kwargs):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        assert len(axes) == len(args), "in_axes must match the positional arguments"
        assert all(a in (0, None) for a in axes)
        n = None
        for a, ax in zip(args, axes):
            if ax == 0:
                n = _np.asarray(_tu.tree_leaves(a)[0]).shape[0]
                break
        if n is None:
            n = _np.asarray(_tu.tree_leaves(kwargs)[0]).shape[0]
        outs = []
        for i in range(n):
            ai = [_tu.tree_map(lambda x: x[i], a) if ax == 0 else a for a, ax in zip(args, axes)]
            ki = {k: _tu.tree_map(lambda x: x[i], v) for k, v in kwargs.items()}
            outs.append(fun(*ai, **ki))
        return _tu.tree_stack(outs)

    return mapped
