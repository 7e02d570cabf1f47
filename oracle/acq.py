"""Restatement of ``/root/reference/bayex/acq.py`` closed forms (TEST INFRASTRUCTURE ONLY).

``norm.cdf`` / ``norm.pdf`` follow jax.scipy (``ndtr`` with the erf/erfc branch
at |x| < sqrt(1/2); pdf = exp(-(log 2pi + x^2)/2)).  All four functions accept
either raw inputs (recomputing the factor like the reference does, quirk Q7) or
a prebuilt ``oracle.gp.Factor`` through ``factor=``.
"""

from __future__ import annotations

import numpy as np
from scipy.special import erf, erfc

from .gp import Factor

ACQ_IDS = {"EI": 0, "PI": 1, "UCB": 2, "LCB": 3}


def ndtr(x):
    dt = x.dtype.type
    h = dt(np.sqrt(0.5))
    w = x * h
    z = np.abs(w)
    with np.errstate(invalid="ignore"):
        y = np.where(z < h, dt(1) + erf(w), np.where(w > 0, dt(2) - erfc(z), erfc(z)))
    return (dt(0.5) * y).astype(x.dtype)


def norm_pdf(x):
    dt = x.dtype.type
    return np.exp(-(dt(np.log(2 * np.pi)) + x * x) / dt(2)).astype(x.dtype)


def _mu_std(x_pred, xs, ys, mask, gp_params, dtype, factor):
    f = factor if factor is not None else Factor(gp_params, xs, ys, mask, dtype)
    return f.predict(x_pred)


def ystar_masked(ys, mask, dtype=np.float32):
    """acq.py:45: max over valid entries, -inf if none."""
    ys = np.asarray(ys, dtype=dtype)
    mb = np.asarray(mask).astype(bool)
    return np.dtype(dtype).type(ys[mb].max()) if mb.any() else np.dtype(dtype).type(-np.inf)


def ystar_unmasked(ys, dtype=np.float32):
    """acq.py:84: ``ys.max()`` over ALL slots, padding zeros included (quirk Q10)."""
    return np.dtype(dtype).type(np.asarray(ys, dtype=dtype).max())


def ei_from(mu, std, ymax, xi=0.01):
    """acq.py:47-50."""
    dt = mu.dtype.type
    a = mu - dt(ymax) - dt(xi)
    z = a / (std + dt(1e-3))
    return (a * ndtr(z) + std * norm_pdf(z)).astype(mu.dtype)


def pi_from(mu, std, ymax, xi=0.01):
    """acq.py:86-87."""
    dt = mu.dtype.type
    return ndtr(((mu - dt(ymax) - dt(xi)) / std).astype(mu.dtype))


def expected_improvement(x_pred, xs, ys, mask, gp_params, xi=0.01, dtype=np.float32, factor=None):
    """acq.py:8-50."""
    ymax = ystar_masked(ys, mask, dtype)
    mu, std = _mu_std(x_pred, xs, ys, mask, gp_params, dtype, factor)
    return ei_from(mu, std, ymax, xi)


def probability_improvement(x_pred, xs, ys, mask, gp_params, xi=0.01, dtype=np.float32, factor=None):
    """acq.py:53-87."""
    y_max = ystar_unmasked(ys, dtype)
    mu, std = _mu_std(x_pred, xs, ys, mask, gp_params, dtype, factor)
    return pi_from(mu, std, y_max, xi)


def upper_confidence_bounds(x_pred, xs, ys, mask, gp_params, kappa=0.01, dtype=np.float32, factor=None):
    """acq.py:90-121."""
    mu, std = _mu_std(x_pred, xs, ys, mask, gp_params, dtype, factor)
    return (mu + mu.dtype.type(kappa) * std).astype(mu.dtype)


def lower_confidence_bounds(x_pred, xs, ys, mask, gp_params, kappa=2.576, dtype=np.float32, factor=None):
    """acq.py:124-156."""
    mu, std = _mu_std(x_pred, xs, ys, mask, gp_params, dtype, factor)
    return (mu - mu.dtype.type(kappa) * std).astype(mu.dtype)


BY_NAME = {"EI": expected_improvement, "PI": probability_improvement,
           "UCB": upper_confidence_bounds, "LCB": lower_confidence_bounds}


def acq_from_moments(name, mu, std, ys, mask, param=None):
    """Closed form on precomputed (mu, std); ``param`` = xi or kappa (reference defaults if None)."""
    dt = mu.dtype
    if name == "EI":
        return ei_from(mu, std, ystar_masked(ys, mask, dt), 0.01 if param is None else param)
    if name == "PI":
        return pi_from(mu, std, ystar_unmasked(ys, dt), 0.01 if param is None else param)
    if name == "UCB":
        return (mu + dt.type(0.01 if param is None else param) * std).astype(dt)
    if name == "LCB":
        return (mu - dt.type(2.576 if param is None else param) * std).astype(dt)
    raise ValueError(f"Acquisition function {name} is not implemented")
