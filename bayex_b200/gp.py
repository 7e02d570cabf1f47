"""``bayex.gp`` surface (reference: bayex/gp.py) on the CUDA library.

Same names, argument order and defaults.  Inputs may be NumPy arrays, lists or torch tensors; results are
host NumPy float32 (the reference returns JAX arrays).  Every function recomputes what it needs from its
arguments, like the reference; the cached-factor fast path lives in ``Optimizer`` / ``engine.Posterior``.
"""
from __future__ import annotations

import ctypes as C
from collections import namedtuple
from functools import partial

import numpy as np
import torch

from . import _lib, engine
from ._lib import lib, check

MASK_VARIANCE = 1e12  # gp.py:10 (kept for API parity; padded rows are compacted away before the kernels run)

GPParams = namedtuple("GPParams", ["noise", "amplitude", "lengthscale"])  # gp.py:12
GPState = namedtuple("GPState", ["params", "momentums", "scales"])  # gp.py:13 (unused by the reference too)


def softplus(x):
    """gp.py:26-27."""
    return np.logaddexp(np.asarray(x, dtype=np.float32), np.float32(0.0))


def cov(x1, x2, mask1, mask2):
    """Masked pairwise squared-exponential matrix (gp.py:20-23), computed by ``bx_cov`` on the GPU."""
    dev = engine.device()
    st = engine.stream(dev)
    a = np.asarray(x1, dtype=np.float32)
    b = np.asarray(x2, dtype=np.float32)
    a = a[:, None] if a.ndim == 1 else a
    b = b[:, None] if b.ndim == 1 else b
    m1 = np.asarray(mask1, dtype=np.float32).reshape(-1)
    m2 = np.asarray(mask2, dtype=np.float32).reshape(-1)
    with engine.on_stream(st, dev):
        ta, tb, t1, t2 = (torch.from_numpy(np.ascontiguousarray(v)).to(dev) for v in (a, b, m1, m2))
        out = torch.empty((a.shape[0], b.shape[0]), dtype=torch.float32, device=dev)
        check(lib.bx_cov(engine._ptr(ta), a.shape[0], engine._ptr(tb), b.shape[0], a.shape[1], engine._ptr(t1),
                         engine._ptr(t2), engine._ptr(out), C.c_void_p(st.cuda_stream)), "bx_cov")
        return out.cpu().numpy()


def exp_quadratic(x1, x2, mask):
    """exp(-|x1 - x2|^2) * mask (gp.py:15-17) - the 1 x 1 case of ``cov``."""
    a = np.atleast_1d(np.asarray(x1, dtype=np.float32)).reshape(1, -1)
    b = np.atleast_1d(np.asarray(x2, dtype=np.float32)).reshape(1, -1)
    return np.float32(cov(a, b, [1.0], [float(mask)])[0, 0])


def _posterior(params, x, y, mask, want_tc=False):
    xv, yv = engine.compact(x, y, mask)
    return engine.Posterior(xv, yv, engine.raw_vector(params, xv.shape[1]), want_tc=want_tc)


def gaussian_process(params, x, y, mask, xt=None, compute_ml=False):
    """gp.py:30-71: NLML (``compute_ml=True``) or posterior (mean, std) at ``xt``."""
    if compute_ml:
        return np.float32(_posterior(params, x, y, mask).nlml_grad()[0])
    assert xt is not None, "xt can't be None during prediction."  # gp.py:59
    xt = np.asarray(xt, dtype=np.float32)
    xt = xt[:, None] if xt.ndim == 1 else xt
    # point sets of BX_POINTS_TC_MIN and more are scored by the tcgen05 engine, which needs the fp16 copies of the factor
    post = _posterior(params, x, y, mask, want_tc=len(xt) >= _lib.POINTS_TC_MIN)
    out = post.score_points(xt, want=("mu", "sigma"))
    with engine.on_stream(post.stream, post.dev):
        return out["mu"].cpu().numpy(), out["sigma"].cpu().numpy()


marginal_likelihood = partial(gaussian_process, compute_ml=True)  # gp.py:74
predict = partial(gaussian_process, compute_ml=False)  # gp.py:76


def grad_fun(params, x, y, mask):
    """d marginal_likelihood / d raw params (gp.py:75), analytic on the GPU; returns GPParams of gradients."""
    post = _posterior(params, x, y, mask)
    g = post.nlml_grad()[3]
    return GPParams(np.float32(g[0]).reshape(np.shape(params.noise)), np.float32(g[1]).reshape(np.shape(params.amplitude)),
                    g[2:].reshape(np.shape(params.lengthscale)) if np.ndim(params.lengthscale) else np.float32(g[2]))


def neg_log_likelihood(params, x, y, mask):
    """gp.py:78-87."""
    return np.float32(_posterior(params, x, y, mask).nlml_grad()[1])


def posterior_fit(y, x, mask, params, lr: float = 1e-3, trainsteps: int = 100):
    """gp.py:90-110: ``trainsteps`` of clip(10) + AdamW(lr) on the raw hypers, entirely on the device."""
    post = _posterior(params, x, y, mask)
    post.fit_hypers(lr=lr, steps=trainsteps)
    return params_from_raw(post.raw_host(), post.d)


def params_from_raw(raw, d):
    raw = np.asarray(raw, dtype=np.float32)
    return GPParams(raw[0:1].reshape(1, 1).copy(), raw[1:2].reshape(1, 1).copy(), raw[2:2 + d].reshape(1, d).copy())


def posterior_fit_restarts(y, x, mask, raws, lr=1e-3, trainsteps=100):
    """``posterior_fit`` (gp.py:90-110) from several starting points - BASELINE config 5.  ``raws`` is (R, d + 2) raw
    hypers ``[noise, amplitude, lengthscale...]``; restarts are sharded over the torch.distributed ranks.  Returns the
    GPParams of the restart with the smallest ``neg_log_likelihood`` plus ``(best index, table)``."""
    from . import fit_sweep
    xv, yv = engine.compact(x, y, mask)
    best, idx, table = fit_sweep.fit_restarts(xv, yv, raws, lr=lr, trainsteps=trainsteps)
    return best, idx, table
