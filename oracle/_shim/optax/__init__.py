"""``optax`` stand-in: clip_by_global_norm, adamw, chain, apply_updates; lbfgs is a zero-update stub.

optax is an unpinned third-party dependency of the reference; these are
restatements of its published algorithms (PARITY UNPINNED).  TEST INFRASTRUCTURE ONLY.
"""
import numpy as _np

from jax import tree_util as _tu


class _GT:
    def __init__(self, init, update):
        self.init, self.update = init, update


def clip_by_global_norm(max_norm):
    def update(g, state, params=None):
        leaves = _tu.tree_leaves(g)
        dt = _np.asarray(leaves[0]).dtype.type
        gn = _np.sqrt(sum(_np.sum(_np.square(_np.asarray(l)), dtype=dt) for l in leaves))
        if gn < max_norm:
            return g, state
        return _tu.tree_map(lambda x: (x / gn) * dt(max_norm), g), state
    return _GT(lambda p: (), update)


def adamw(learning_rate, b1=0.9, b2=0.999, eps=1e-8, weight_decay=1e-4):
    def init(p):
        z = lambda x: _np.zeros_like(_np.asarray(x))
        return dict(mu=_tu.tree_map(z, p), nu=_tu.tree_map(z, p), count=0)

    def update(g, st, params=None):
        dt = _np.asarray(_tu.tree_leaves(g)[0]).dtype.type
        mu = _tu.tree_map(lambda m, x: dt(b1) * m + dt(1 - b1) * x, st["mu"], g)
        nu = _tu.tree_map(lambda v, x: dt(b2) * v + dt(1 - b2) * x * x, st["nu"], g)
        t = st["count"] + 1
        c1, c2 = dt(1 - dt(b1) ** t), dt(1 - dt(b2) ** t)
        upd = _tu.tree_map(lambda m, v, p: -dt(learning_rate) * ((m / c1) / (_np.sqrt(v / c2) + dt(eps))
                                                                    + dt(weight_decay) * _np.asarray(p)),
                           mu, nu, params)
        return upd, dict(mu=mu, nu=nu, count=t)
    return _GT(init, update)


def chain(*ts):
    def init(p):
        return [t.init(p) for t in ts]

    def update(g, states, params=None):
        new = []
        for t, s in zip(ts, states):
            g, s = t.update(g, s, params)
            new.append(s)
        return g, new
    return _GT(init, update)


def apply_updates(params, updates):
    return _tu.tree_map(lambda p, u: (_np.asarray(p) + u).astype(_np.asarray(p).dtype), params, updates)


def lbfgs():
    """Zero-update stub: the reference's refinement step becomes the identity under the shim."""
    def update(g, state, params=None, **_kw):
        return _tu.tree_map(lambda x: _np.zeros_like(_np.asarray(x)), g), state
    return _GT(lambda p: (), update)


def value_and_grad_from_state(fun):
    def vg(params, state=None):
        return fun(params), _np.zeros_like(_np.asarray(params))
    return vg
