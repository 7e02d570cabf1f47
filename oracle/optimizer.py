"""Restatement of ``/root/reference/bayex/optimizer.py`` on NumPy (TEST INFRASTRUCTURE ONLY).

State machine (init / sample / expand / fit) follows the reference line by
line.  Two documented differences, both from SURVEY quirks:

* ``sample`` scores candidates with the diagonal-only posterior (quirk Q8; the
  reference's M x M product is infeasible at the benchmark sizes and its
  diagonal is mathematically identical).
* The L-BFGS refinement (``optimizer.py:39-73``) lives in optax (unpinned, and
  fed a mis-signed ``value_fn`` - quirk Q12), so its trajectory is PARITY
  UNPINNED.  ``refine="lbfgs"`` / ``"lbfgs_quirk"`` run a restatement of the
  published algorithm (``oracle/lbfgs.py``: L-BFGS two-loop recursion + strong-
  Wolfe zoom search, not optax's code); ``refine="identity"`` skips it (what the golden vectors made from
  the reference source under the shim use, where ``optax.lbfgs`` is stubbed to
  a zero update); ``refine="ascent"`` restates the monotone multi-start ascent
  the CUDA path runs (``bayex_b200/csrc/refine.cu``) so the contract
  "refined value >= start value, proposal value >= best random candidate" can
  be checked on identical inputs.
"""

from __future__ import annotations

from typing import NamedTuple

import numpy as np

from . import acq as oacq
from .domain import ParamSpace
from .gp import Factor, GPParams, posterior_fit

N_STARTS = 50  # optimizer.py:196
REFINE_ROUNDS = 30


class OptimizerState(NamedTuple):
    """optimizer.py:14-36."""
    params: dict
    ys: np.ndarray
    best_score: float
    best_params: dict
    mask: np.ndarray
    gp_params: GPParams


def acq_value_and_grad(name, f: Factor, ys, mask, x, param=None):
    """Acquisition value and analytic gradient at points x (S, d).

    Replaces the autodiff of optimizer.py:60-64.  Uses K^-1 k = L^-T (L^-1 k).
    """
    from scipy.linalg import solve_triangular
    dt = f.dtype
    x = np.asarray(x, dtype=dt)
    u = x / f.ls
    S, d = u.shape
    diff = f.xs[None, :, :] - u[:, None, :]  # (S, n, d): u_j - u
    k = f.amp * np.exp(-np.sum(diff * diff, axis=-1)) * f.m[None, :]  # (S, n)
    g = dt(2.0) * diff / f.ls[None, None, :]  # d k_j / d x_k = k_j * g_jk
    mu = k @ (f.alpha * f.m) + f.ymean
    v = solve_triangular(f.L, k.T, lower=True, check_finite=False)  # (n, S)
    w = solve_triangular(f.L.T, v, lower=False, check_finite=False).T  # (S, n)
    var = f.amp - np.sum(v * v, axis=0)
    dmu = np.einsum("sn,snk->sk", k * (f.alpha * f.m)[None, :], g)
    dvar = -dt(2.0) * np.einsum("sn,snk->sk", w * k, g)
    std = np.sqrt(np.maximum(var, dt(1e-10)))
    dstd = np.where((var > dt(1e-10))[:, None], dvar / (dt(2.0) * std[:, None]), dt(0.0))
    if name == "EI":
        xi = dt(0.01 if param is None else param)
        A = mu - oacq.ystar_masked(ys, mask, dt) - xi
        sp = std + dt(1e-3)
        z = A / sp
        cdf, pdf = oacq.ndtr(z), oacq.norm_pdf(z)
        val = A * cdf + std * pdf
        dz = dmu / sp[:, None] - (A / (sp * sp))[:, None] * dstd
        grad = cdf[:, None] * dmu + ((A - std * z) * pdf)[:, None] * dz + pdf[:, None] * dstd
    elif name == "PI":
        xi = dt(0.01 if param is None else param)
        A = mu - oacq.ystar_unmasked(ys, dt) - xi
        z = A / std
        val = oacq.ndtr(z)
        grad = oacq.norm_pdf(z)[:, None] * (dmu / std[:, None] - (A / (std * std))[:, None] * dstd)
    elif name == "UCB":
        kap = dt(0.01 if param is None else param)
        val, grad = mu + kap * std, dmu + kap * dstd
    elif name == "LCB":
        kap = dt(2.576 if param is None else param)
        val, grad = mu - kap * std, dmu - kap * dstd
    else:
        raise ValueError(f"Acquisition function {name} is not implemented")
    return val.astype(dt), grad.astype(dt)


def refine_ascent(name, f: Factor, ys, mask, starts, rounds=REFINE_ROUNDS, param=None):
    """Monotone multi-start ascent (the CUDA ``bx_refine`` algorithm, restated).

    Per start: trial = best + s * (g*ls^2)/||g*ls||, accepted iff the value
    strictly improves (s doubles) else rejected (s halves); s0 = 0.1 (in
    length-scale units); iterates clipped to +-1e6 like optimizer.py:68.
    """
    dt = f.dtype
    xb = np.asarray(starts, dtype=dt).copy()
    fb, gb = acq_value_and_grad(name, f, ys, mask, xb, param)
    step = np.full((xb.shape[0],), dt(0.1), dtype=dt)
    for _ in range(rounds):
        gl = gb * f.ls[None, :]
        nrm = np.sqrt(np.sum(gl * gl, axis=1))
        ok = np.isfinite(nrm) & (nrm > 0)
        direction = np.where(ok[:, None], gl * f.ls[None, :] / np.where(ok, nrm, 1)[:, None], 0)
        xt = np.clip(xb + step[:, None] * direction, dt(-1e6), dt(1e6)).astype(dt)
        ft, gt = acq_value_and_grad(name, f, ys, mask, xt, param)
        acc = ft > fb
        xb = np.where(acc[:, None], xt, xb)
        gb = np.where(acc[:, None], gt, gb)
        fb = np.where(acc, ft, fb)
        step = np.where(acc, step * dt(2.0), step * dt(0.5)).astype(dt)
    return xb, fb


def jnp_argmax(v):
    """jnp.argmax: first occurrence; NaN counts as the maximum (first NaN wins)."""
    v = np.asarray(v)
    nan = np.isnan(v)
    return int(np.argmax(nan)) if nan.any() else int(np.argmax(v))


def jnp_argsort(v):
    """jnp.argsort: stable ascending, NaNs last, -0.0 == +0.0."""
    return np.argsort(np.asarray(v) + 0.0, kind="stable")


class Optimizer:
    """optimizer.py:76-280."""

    def __init__(self, domain: dict, acq: str = "EI", maximize: bool = False, dtype=np.float32):
        self.domain = domain
        self.sign = 1 if maximize else -1  # optimizer.py:95
        self.param_space = ParamSpace(domain)
        if acq not in oacq.BY_NAME:
            raise ValueError(f"Acquisition function {acq} is not implemented")  # optimizer.py:107
        self.acq_name = acq
        self.acq = oacq.BY_NAME[acq]
        self.dtype = dtype

    def init(self, ys, params: dict, noise_scale: float = -8.0, fit: bool = True):
        """optimizer.py:109-166."""
        num_entries = len(ys)
        pad_value = int(np.ceil(len(ys) / 10) * 10)  # :123
        ys = self.sign * np.asarray(ys, dtype=np.float32).reshape(-1)  # :126-127
        mask = np.zeros((pad_value,), dtype=bool)
        mask[:num_entries] = True  # :131
        ysp = np.zeros((pad_value,), dtype=ys.dtype)
        ysp[:num_entries] = ys  # :132
        _params = {}
        for key, entries in params.items():
            assert key in self.domain, f"Parameter {key} is not in the domain"  # :137
            values = np.zeros((pad_value,), dtype=self.domain[key].dtype)  # :140 (int32 for Integer)
            values[:num_entries] = np.asarray(entries)
            _params[key] = values
        best_score = float(np.max(ysp[mask]))  # :146
        best_idx = int(np.argmax(ysp[mask]))
        best_params = {k: v[mask][best_idx] for k, v in _params.items()}  # :148
        d = len(_params)
        gp_params = GPParams(np.full((1, 1), 1.0 * noise_scale, dtype=np.float32),
                             np.zeros((1, 1), dtype=np.float32),
                             np.zeros((1, d), dtype=np.float32))  # :151-155
        xs = np.stack([self.domain[k].transform(_params[k]) for k in _params], axis=1)  # :158 (params order, Q9)
        if fit:
            gp_params = posterior_fit(ysp, xs, mask, gp_params, dtype=self.dtype)  # :159
        return OptimizerState(_params, self.sign * ysp, self.sign * best_score, best_params, mask, gp_params)

    def sample(self, key, state, size=10_000, refine="identity", n_starts=N_STARTS, details=False, offset=0):
        """optimizer.py:168-208."""
        samples = self.param_space.sample_params(key, (size,), offset=offset)  # :183
        xs_samples = self.param_space.to_array(samples)  # :184
        xs = self.param_space.to_array(state.params)  # :187
        ys = (self.sign * state.ys).astype(np.float32)  # :188
        f = Factor(state.gp_params, xs, ys, state.mask, self.dtype)
        acq_vals = self.acq(xs_samples, xs, ys, state.mask, state.gp_params, dtype=self.dtype, factor=f)  # :193
        top_idxs = jnp_argsort(acq_vals)[-n_starts:]  # :196
        init_points = xs_samples[top_idxs]  # :197
        if refine == "identity":
            optimized = init_points.astype(self.dtype)
        elif refine == "ascent":
            optimized, _ = refine_ascent(self.acq_name, f, ys, state.mask, init_points)
        elif refine in ("lbfgs", "lbfgs_quirk"):
            # optimizer.py:198-199: vmap(_optimize_suggestion)(init_points) with max_iter = 10 - L-BFGS restated from the
            # published algorithm (oracle/lbfgs.py, PARITY UNPINNED); "lbfgs_quirk" wires the signs as the reference does (Q12)
            from . import lbfgs as olbfgs
            f64 = Factor(state.gp_params, xs, ys, state.mask, np.float64)

            def fg(x):
                v, g = acq_value_and_grad(self.acq_name, f64, ys.astype(np.float64), state.mask, x[None, :])
                return float(v[0]), g[0].astype(np.float64)

            optimized = np.stack([olbfgs.optimize_suggestion(p0.astype(np.float64), fg, 10, sign_quirk=(refine == "lbfgs_quirk"))
                                  for p0 in init_points]).astype(self.dtype)
        else:
            raise ValueError(refine)
        opt_vals = self.acq(optimized, xs, ys, state.mask, state.gp_params, dtype=self.dtype, factor=f)  # :200
        all_points = np.concatenate((optimized, xs_samples.astype(self.dtype)), axis=0)  # :203
        all_vals = np.concatenate((opt_vals, acq_vals), axis=0)  # :204
        chosen = jnp_argmax(all_vals)  # :205
        best = self.param_space.to_dict(all_points[chosen][None])  # :207
        best = {k: np.asarray(v, dtype=np.float32) for k, v in best.items()}
        if details:
            return best, dict(xs_samples=xs_samples, acq_vals=acq_vals, top_idxs=top_idxs, optimized=optimized,
                              opt_vals=opt_vals, chosen=chosen, all_vals=all_vals, factor=f)
        return best

    def expand(self, st: OptimizerState):
        """optimizer.py:211-239."""
        if int(np.sum(st.mask)) == len(st.mask):
            pad_value = int(np.ceil(len(st.mask) * 2 / 10) * 10)  # :224
            diff = pad_value - len(st.mask)
            mask = np.pad(st.mask, (0, diff))
            ys = np.pad(st.ys, (0, diff))
            params = {k: np.pad(v, (0, diff)) for k, v in st.params.items()}
        else:
            mask, ys, params = st.mask, st.ys, st.params
        return OptimizerState(params, ys, st.best_score, st.best_params, mask, st.gp_params)

    def fit(self, st: OptimizerState, y, new_params, fit: bool = True):
        """optimizer.py:242-280."""
        st = self.expand(st)
        slot = int(np.argmin(st.mask))  # :261 first False
        last = np.arange(len(st.mask)) == slot
        mask = np.where(last, True, st.mask)
        ys = np.where(last, np.float32(np.asarray(y).reshape(-1)[0]), st.ys).astype(np.float32)  # :263
        params = {}
        for k, v in st.params.items():
            nv = np.asarray(new_params[k]).reshape(-1)[0]
            params[k] = np.where(last, nv, v)  # :264 (promotes int32 storage to float32, quirk Q14)
            if params[k].dtype == np.float64:
                params[k] = params[k].astype(np.float32)
        xs = np.stack([self.domain[k].transform(params[k]) for k in params], axis=1)  # :266-267
        ys_s = (self.sign * ys).astype(np.float32)
        gp_params = posterior_fit(ys_s, xs, mask, st.gp_params, dtype=self.dtype) if fit else st.gp_params  # :269
        best_score = np.max(ys_s[mask])  # :271
        best_idx = int(np.argmax(np.where(mask, ys_s, -np.inf)))  # :272
        best_params = {k: v[best_idx] for k, v in params.items()}
        return OptimizerState(params, ys, self.sign * best_score, best_params, mask, gp_params)
