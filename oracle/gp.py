"""NumPy/SciPy restatement of ``/root/reference/bayex/gp.py`` (TEST INFRASTRUCTURE ONLY).

Every function cites the reference lines it follows.  ``dtype`` selects the
arithmetic: float32 mirrors the reference (JAX default, x64 off), float64 is
the arbiter used to judge fp32 round-off.  The padded/masked formulation of the
reference is kept verbatim here (the CUDA path compacts to the valid rows - the
two are equivalent to ~1e-12, reference tests ``tests/test_gp.py:54-99``).
"""

from __future__ import annotations

from collections import namedtuple

import numpy as np
from scipy.linalg import cholesky, solve_triangular

MASK_VARIANCE = 1e12  # gp.py:10

GPParams = namedtuple("GPParams", ["noise", "amplitude", "lengthscale"])  # gp.py:12
GPState = namedtuple("GPState", ["params", "momentums", "scales"])  # gp.py:13 (dead type)

PRIORS = GPParams(-8.0, 1.0, 1.0)  # gp.py:83


def softplus(x):
    """gp.py:26-27."""
    return np.logaddexp(x, 0.0)


def exp_quadratic(x1, x2, mask):
    """gp.py:15-17: exp(-sum((x1-x2)^2)) * mask.  No 1/2 factor (quirk Q1)."""
    return np.exp(-np.sum((np.asarray(x1) - np.asarray(x2)) ** 2)) * mask


def _as2d(x, dtype):
    x = np.asarray(x, dtype=dtype)
    return x[:, None] if x.ndim == 1 else x


def cov(x1, x2, mask1, mask2, dtype=np.float32):
    """gp.py:20-23: pairwise exp_quadratic with outer-product mask, (n1, n2)."""
    a = _as2d(x1, dtype)
    b = _as2d(x2, dtype)
    m = np.outer(np.asarray(mask1, dtype=dtype), np.asarray(mask2, dtype=dtype))
    diff = a[:, None, :] - b[None, :, :]
    d2 = np.sum(diff * diff, axis=-1, dtype=dtype)
    return (np.exp(-d2) * m).astype(dtype)


def _cov_cols(a, b, mask1, dtype):
    """cov(a, b, mask1, ones) without the (n1,n2,d) temporary (large n / M)."""
    out = np.zeros((a.shape[0], b.shape[0]), dtype=dtype)
    for k in range(a.shape[1]):
        diff = a[:, k][:, None] - b[:, k][None, :]
        out += diff * diff
    np.negative(out, out=out)
    np.exp(out, out=out)
    out *= mask1[:, None]
    return out


def unpack_params(params, d, dtype):
    """Raw GPParams (floats or (1,1)/(1,d) arrays, quirk Q2) -> (noise_raw, amp_raw, ls_raw[d])."""
    nr = dtype(np.asarray(params.noise, dtype=dtype).reshape(-1)[0])
    ar = dtype(np.asarray(params.amplitude, dtype=dtype).reshape(-1)[0])
    ls = np.asarray(params.lengthscale, dtype=dtype).reshape(-1)
    if ls.size == 1 and d != 1:
        ls = np.full((d,), ls[0], dtype=dtype)
    assert ls.size == d, "lengthscale must broadcast against the input dimension"
    return nr, ar, ls.astype(dtype)


class Factor:
    """Everything the predict half needs, computed once (what gp.py:41-49 recomputes per call)."""

    def __init__(self, params, x, y, mask, dtype=np.float32):
        dtype = np.dtype(dtype).type
        x = _as2d(x, dtype)
        n, d = x.shape
        mask_b = np.asarray(mask).astype(bool)
        m = mask_b.astype(dtype)
        y = np.asarray(y, dtype=dtype)
        nr, ar, lr = unpack_params(params, d, dtype)
        self.dtype = dtype
        self.noise, self.amp, self.ls = dtype(softplus(nr)), dtype(softplus(ar)), softplus(lr).astype(dtype)  # gp.py:41
        self.ymean = dtype(np.mean(y[mask_b], dtype=dtype))  # gp.py:43
        self.yc = ((y - self.ymean) * m).astype(dtype)  # gp.py:44
        self.xs = (x / self.ls).astype(dtype)  # gp.py:45
        self.m = m
        self.n = n
        K = self.amp * (_cov_cols(self.xs, self.xs, m, dtype) * m[None, :])  # gp.py:46 (cov with outer mask)
        K = K + np.eye(n, dtype=dtype) * dtype(self.noise + dtype(1e-6))  # gp.py:46
        K = K + np.eye(n, dtype=dtype) * (dtype(1.0) - m) * dtype(MASK_VARIANCE)  # gp.py:47
        self.K = K.astype(dtype)
        self.L = cholesky(self.K, lower=True, check_finite=False).astype(dtype)  # gp.py:48
        t = solve_triangular(self.L, self.yc, lower=True, check_finite=False)
        self.alpha = solve_triangular(self.L.T, t, lower=False, check_finite=False).astype(dtype)  # gp.py:49

    def nlml(self):
        """gp.py:51-57: the *negative* log marginal likelihood plus the amplitude term (quirk Q4)."""
        dt = self.dtype
        logp = dt(0.5) * np.dot(self.yc, self.alpha)  # :52
        logp += np.sum(np.log(np.diag(self.L)), dtype=dt)  # :53
        logp -= np.sum(dt(1.0) - self.m, dtype=dt) * dt(0.5) * dt(np.log(MASK_VARIANCE))  # :54
        logp += (np.sum(self.m, dtype=dt) / dt(2)) * dt(np.log(2 * np.pi))  # :55
        logp += dt(-0.5 * np.log(2 * np.pi)) - np.log(self.amp) - np.log(self.amp) ** 2  # :56
        return dt(logp)

    def predict(self, xt, chunk=8192, return_var=False, threads=1):
        """gp.py:59-71 with the diagonal of pred_var only (quirk Q8), chunked over candidates.

        ``threads > 1`` (bench.py's CPU arm only): the chunks are independent, so they are spread over a thread pool -
        NumPy releases the GIL inside the element-wise Gram build and inside BLAS; the caller limits BLAS itself to one
        thread per call (threadpoolctl) so that ``threads`` cores are busy through both stages."""
        dt = self.dtype
        xt = _as2d(xt, dt)
        xt = (xt / self.ls).astype(dt)  # gp.py:60
        M = xt.shape[0]
        mean = np.empty((M,), dtype=dt)
        var = np.empty((M,), dtype=dt)
        am = (self.alpha * self.m).astype(dt)  # gp.py:66

        def one(s):
            kc = self.amp * _cov_cols(self.xs, xt[s:s + chunk], self.m, dt)  # gp.py:64
            mean[s:s + chunk] = kc.T @ am + self.ymean  # gp.py:67
            v = solve_triangular(self.L, kc, lower=True, check_finite=False)  # gp.py:68
            # diag(amp*cov(xt,xt) - v.T@v): cov(xt,xt)_jj == 1  (gp.py:69-70)
            var[s:s + chunk] = self.amp - np.sum(v * v, axis=0, dtype=dt)

        if threads > 1:
            from concurrent.futures import ThreadPoolExecutor
            with ThreadPoolExecutor(max_workers=threads) as pool:
                list(pool.map(one, range(0, M, chunk)))
        else:
            for s in range(0, M, chunk):
                one(s)
        std = np.sqrt(np.maximum(var, dt(1e-10))).astype(dt)  # gp.py:70
        return (mean, std, var) if return_var else (mean, std)


def gaussian_process(params, x, y, mask, xt=None, compute_ml=False, dtype=np.float32):
    """gp.py:30-71."""
    f = Factor(params, x, y, mask, dtype)
    if compute_ml:
        return f.nlml()
    assert xt is not None, "xt can't be None during prediction."  # gp.py:59
    return f.predict(xt)


def marginal_likelihood(params, x, y, mask, dtype=np.float32):
    """gp.py:74."""
    return gaussian_process(params, x, y, mask, compute_ml=True, dtype=dtype)


def predict(params, x, y, mask, xt=None, dtype=np.float32):
    """gp.py:76."""
    return gaussian_process(params, x, y, mask, xt=xt, compute_ml=False, dtype=dtype)


def predict_full_cov(params, x, y, mask, xt, dtype=np.float32):
    """gp.py:63-71 exactly as written, *including* the M x M product (small M only)."""
    f = Factor(params, x, y, mask, dtype)
    dt = f.dtype
    xt = (_as2d(xt, dt) / f.ls).astype(dt)
    ones = np.ones(len(xt), dtype=dt)
    kc = f.amp * cov(f.xs, xt, f.m, ones, dt)
    mean = kc.T @ (f.alpha * f.m) + f.ymean
    v = solve_triangular(f.L, kc, lower=True, check_finite=False)
    pv = f.amp * cov(xt, xt, ones, ones, dt) - v.T @ v
    return mean.astype(dt), np.sqrt(np.maximum(np.diag(pv), dt(1e-10))).astype(dt)


def neg_log_likelihood(params, x, y, mask, dtype=np.float32):
    """gp.py:78-87: -(NLML - 1/2 sum (raw - prior)^2)  (sign-inverted data term, quirk Q5)."""
    dt = np.dtype(dtype).type
    ll = marginal_likelihood(params, x, y, mask, dtype)
    d = _as2d(x, dt).shape[1]
    nr, ar, lr = unpack_params(params, d, dt)
    log_prior = (nr - dt(PRIORS.noise)) ** 2 + (ar - dt(PRIORS.amplitude)) ** 2 + np.sum(
        (lr - dt(PRIORS.lengthscale)) ** 2, dtype=dt)
    return dt(-(ll - dt(0.5) * log_prior))


def nlml_grad(params, x, y, mask, dtype=np.float32):
    """Analytic d(NLML)/d(raw params) = what ``grad_fun`` (gp.py:75) returns via autodiff.

    With W = K^-1 - alpha alpha^T:  dNLML/dtheta = 1/2 tr(W dK/dtheta); plus the
    amplitude term of gp.py:56; chain rule through softplus (sigmoid).
    Returns GPParams of raw-parameter gradients (noise, amplitude, lengthscale[d]).
    """
    f = Factor(params, x, y, mask, dtype)
    dt = f.dtype
    d = f.xs.shape[1]
    nr, ar, lr = unpack_params(params, d, dt)
    n = f.n
    Tm = solve_triangular(f.L, np.eye(n, dtype=dt), lower=True, check_finite=False)
    Kinv = (Tm.T @ Tm).astype(dt)
    W = Kinv - np.outer(f.alpha, f.alpha)
    C = _cov_cols(f.xs, f.xs, f.m, dt) * f.m[None, :]
    sig = lambda r: dt(1.0) / (dt(1.0) + np.exp(-r))
    g_noise = dt(0.5) * np.trace(W) * sig(nr)
    la = np.log(f.amp)
    g_amp = (dt(0.5) * np.sum(W * C, dtype=dt) - dt(1.0) / f.amp - dt(2.0) * la / f.amp) * sig(ar)
    WC = W * C * f.amp
    g_ls = np.empty((d,), dtype=dt)
    for k in range(d):
        diff = f.xs[:, k][:, None] - f.xs[:, k][None, :]  # already divided by ls
        # d/dls exp(-(dx/ls)^2) = exp(.) * 2 dx^2 / ls^3 = exp(.) * 2 (dx/ls)^2 / ls
        g_ls[k] = dt(0.5) * np.sum(WC * dt(2.0) * diff * diff, dtype=dt) / f.ls[k]
    g_ls = g_ls * sig(lr)
    return GPParams(dt(g_noise), dt(g_amp), g_ls.astype(dt))


def loss_grad(params, x, y, mask, dtype=np.float32):
    """Gradient of ``neg_log_likelihood`` (gp.py:78-87) w.r.t. the raw params (what gp.py:104 computes)."""
    dt = np.dtype(dtype).type
    g = nlml_grad(params, x, y, mask, dtype)
    d = _as2d(x, dt).shape[1]
    nr, ar, lr = unpack_params(params, d, dt)
    return GPParams(dt(-g.noise + (nr - dt(PRIORS.noise))),
                    dt(-g.amplitude + (ar - dt(PRIORS.amplitude))),
                    (-g.lengthscale + (lr - dt(PRIORS.lengthscale))).astype(dt))


def adamw_step(p, g, m, v, t, lr, dt, b1=0.9, b2=0.999, eps=1e-8, wd=1e-4, max_norm=10.0):
    """One ``chain(clip_by_global_norm(10), adamw(lr))`` update on a flat vector (gp.py:99,105-106).

    optax is an unpinned dependency; restated from its published algorithm - PARITY UNPINNED.
    """
    gn = np.sqrt(np.sum(g * g, dtype=dt))
    g = g if gn < dt(max_norm) else (g / gn * dt(max_norm)).astype(dt)
    m = (dt(b1) * m + dt(1 - b1) * g).astype(dt)
    v = (dt(b2) * v + dt(1 - b2) * g * g).astype(dt)
    t = t + 1
    mh = m / dt(1 - dt(b1) ** t)
    vh = v / dt(1 - dt(b2) ** t)
    u = mh / (np.sqrt(vh) + dt(eps)) + dt(wd) * p
    p = (p - dt(lr) * u).astype(dt)
    return p, m, v, t


def posterior_fit(y, x, mask, params, lr=1e-3, trainsteps=100, dtype=np.float32, return_trace=False):
    """gp.py:90-110: ``trainsteps`` x {grad of neg_log_likelihood, clip 10, AdamW}; fresh Adam state."""
    dt = np.dtype(dtype).type
    d = _as2d(x, dt).shape[1]
    nr, ar, lr_raw = unpack_params(params, d, dt)
    p = np.concatenate([[nr], [ar], lr_raw]).astype(dt)
    m = np.zeros_like(p)
    v = np.zeros_like(p)
    t = 0
    trace = []
    for _ in range(trainsteps):
        cur = GPParams(p[0], p[1], p[2:])
        g = loss_grad(cur, x, y, mask, dtype)
        gv = np.concatenate([[g.noise], [g.amplitude], g.lengthscale]).astype(dt)
        if return_trace:
            trace.append((p.copy(), gv.copy()))
        p, m, v, t = adamw_step(p, gv, m, v, t, lr, dt)
    out = GPParams(np.full((1, 1), p[0], dtype=dt), np.full((1, 1), p[1], dtype=dt), p[2:].reshape(1, d).astype(dt))
    return (out, trace) if return_trace else out
