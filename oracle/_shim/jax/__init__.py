"""NumPy stand-in for the ``jax`` namespace (see oracle/_shim/README.md). TEST INFRASTRUCTURE ONLY."""
import functools as _ft

import numpy as _np

from . import numpy, lax, random, scipy, tree_util  # noqa: F401
from . import tree_util as tree  # jax.tree.map / jax.tree.leaves
from .numpy import _cfg, _wrap

Array = _np.ndarray


class _Config:
    def update(self, name, value):
        if name == "jax_enable_x64":
            _cfg["float"] = _np.float64 if value else _np.float32
        else:
            raise KeyError(name)


config = _Config()


def jit(fun=None, **_kw):
    if fun is None:
        return lambda f: f
    return fun


def vmap(fun, in_axes=0, out_axes=0):
    def mapped(*args):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        n = next(_np.shape(a)[ax] for a, ax in zip(args, axes) if ax is not None)
        outs = []
        for i in range(n):
            outs.append(fun(*[a if ax is None else _np.take(a, i, axis=ax) for a, ax in zip(args, axes)]))
        return tree_util.tree_map(lambda *xs: _wrap(_np.stack([_np.asarray(x) for x in xs], axis=out_axes)), *outs)
    return mapped


def grad(fun, argnums=0):
    """Central finite differences over the pytree leaves of the first argument.

    The objective is always evaluated in float64 (x64 switched on for the call) so the
    difference quotient is accurate to ~1e-9; the result is cast back to the leaf dtype.
    """
    def g(params, *args, **kw):
        saved = _cfg["float"]
        _cfg["float"] = _np.float64
        try:
            return _g(params, *args, **kw)
        finally:
            _cfg["float"] = saved

    def _g(params, *args, **kw):
        leaves, rebuild = tree_util.tree_flatten(params)
        dts = [_np.asarray(l).dtype if _np.asarray(l).dtype.kind == "f" else _np.dtype(saved_float()) for l in leaves]
        args = tuple(_np.asarray(a, dtype=_np.float64) if isinstance(a, _np.ndarray) and a.dtype.kind == "f" else a
                     for a in args)
        base = [_np.array(l, dtype=_np.float64) for l in leaves]
        outs = []
        for li, leaf in enumerate(base):
            gl = _np.zeros_like(leaf)
            it = _np.nditer(leaf, flags=["multi_index"])
            for _ in it:
                idx = it.multi_index
                h = 1e-6 * max(1.0, abs(float(leaf[idx])))
                vals = []
                for s in (+1, -1):
                    pert = [b.copy() for b in base]
                    pert[li][idx] += s * h
                    vals.append(float(_np.asarray(fun(rebuild(pert), *args, **kw))))
                gl[idx] = (vals[0] - vals[1]) / (2 * h)
            outs.append(gl.astype(dts[li]))
        return rebuild(outs)

    def saved_float():
        return _np.float32
    return g


def value_and_grad(fun, argnums=0):
    gf = grad(fun, argnums)
    return lambda *a, **k: (fun(*a, **k), gf(*a, **k))
