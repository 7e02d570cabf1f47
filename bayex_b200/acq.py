"""``bayex.acq`` names and signatures (reference: bayex/acq.py) on the CUDA scoring kernel.

Each call rebuilds the factor from its arguments exactly like the reference does (quirk Q7); ``Optimizer.sample``
uses the cached factor instead.  Results are host NumPy float32 arrays of shape ``(len(x_pred),)``.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib, engine
from .gp import GPParams  # noqa: F401  (re-exported like the reference: acq.py:5)


def ystar_for(acq: str, ys, mask) -> float:
    """EI: masked max (acq.py:45).  PI: ``ys.max()`` over ALL slots, padding zeros included (acq.py:84, quirk Q10)."""
    ys = np.asarray(ys, dtype=np.float32).reshape(-1)
    if acq == "EI":
        m = np.asarray(mask).astype(bool).reshape(-1)
        return float(ys[m].max()) if m.any() else float("-inf")
    if acq == "PI":
        return float(ys.max())
    return 0.0


def _score(acq, x_pred, xs, ys, mask, gp_params, param):
    xv, yv = engine.compact(xs, ys, mask)
    xt = np.asarray(x_pred, dtype=np.float32)
    xt = xt[:, None] if xt.ndim == 1 else xt
    # large point sets run on the tcgen05 engine (bx_score_points switches at BX_POINTS_TC_MIN), which needs the fp16
    # copies of the factor; small ones stay on the fp32 SIMT engine and skip building them
    post = engine.Posterior(xv, yv, engine.raw_vector(gp_params, xv.shape[1]), want_tc=len(xt) >= _lib.POINTS_TC_MIN)
    out = post.score_points(xt, acq=acq, param=param, ystar=ystar_for(acq, ys, mask), want=("acq",))
    with engine.on_stream(post.stream, post.dev):
        return out["acq"].cpu().numpy()


def expected_improvement(x_pred, xs, ys, mask, gp_params, xi: float = 0.01):
    """EI = (mu - y* - xi) Phi(z) + sigma phi(z), z = (mu - y* - xi) / (sigma + 1e-3)  (acq.py:8-50)."""
    return _score("EI", x_pred, xs, ys, mask, gp_params, xi)


def probability_improvement(x_pred, xs, ys, mask, gp_params, xi: float = 0.01):
    """PI = Phi((mu - y* - xi) / sigma)  (acq.py:53-87)."""
    return _score("PI", x_pred, xs, ys, mask, gp_params, xi)


def upper_confidence_bounds(x_pred, xs, ys, mask, gp_params, kappa: float = 0.01):
    """UCB = mu + kappa sigma  (acq.py:90-121)."""
    return _score("UCB", x_pred, xs, ys, mask, gp_params, kappa)


def lower_confidence_bounds(x_pred, xs, ys, mask, gp_params, kappa: float = 2.576):
    """LCB = mu - kappa sigma  (acq.py:124-156)."""
    return _score("LCB", x_pred, xs, ys, mask, gp_params, kappa)


DEFAULT_PARAM = {"EI": 0.01, "PI": 0.01, "UCB": 0.01, "LCB": 2.576}  # acq.py:14,59,96,130
