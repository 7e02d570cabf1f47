"""L-BFGS + strong-Wolfe zoom line search on NumPy (TEST INFRASTRUCTURE ONLY) - SURVEY 8(f) rank 2.

The reference refines its 50 start points with ``optax.lbfgs()`` (``optimizer.py:39-73``).  optax is an unpinned third-party
dependency whose source is not available here, and the reference's own tests assert nothing about the refined points, so
this restatement is **PARITY UNPINNED**: it follows the *published* algorithms optax documents itself as implementing -
Nocedal & Wright, "Numerical Optimization" (2nd ed.), Algorithm 7.4 (two-loop recursion) with the initial scaling
gamma = s^T y / y^T y (7.20), memory 10, and Algorithms 3.5 / 3.6 (line search for the strong Wolfe conditions with zoom),
c1 = 1e-4, c2 = 0.9, at most 20 function evaluations per search, first trial step 1 - not optax's code.  It exists so that
the role the refinement plays in ``sample()`` can be exercised with the reference's own call pattern, including the sign
quirk Q12 of ``optimizer.py:59-66``:

    value_and_grad_fun = optax.value_and_grad_from_state(lambda x: -fun(x))
    updates, state = opt.update(grad, state, params, value=value, grad=grad, value_fn=fun)      # value_fn is NOT negated

i.e. the search direction is built from the gradient of ``-fun`` (descent on -fun = ascent on fun) while the line search
evaluates ``+fun`` along it and (per optax's ``value_and_grad_from_state``) its stored value / gradient are reused by the
next step.  ``sign_quirk=True`` reproduces that wiring; ``False`` is the refinement the reference's comment describes.

The CUDA path does NOT use this (``bayex_b200/csrc/refine.cu`` is a monotone analytic-gradient ascent whose contract is
value-based and judged by the oracle); nothing under ``bayex_b200/`` imports it.
"""
from __future__ import annotations

import numpy as np

MEMORY = 10
C1, C2 = 1e-4, 0.9
MAX_LS_EVALS = 20


def _two_loop(g, S, Y):
    """Nocedal & Wright Algorithm 7.4: r = H g with H the L-BFGS inverse-Hessian approximation from the pairs (S, Y)."""
    q = g.copy()
    alphas = []
    for s, y in zip(reversed(S), reversed(Y)):
        rho = 1.0 / np.dot(y, s)
        a = rho * np.dot(s, q)
        alphas.append((a, rho, s, y))
        q = q - a * y
    if S:
        gamma = np.dot(S[-1], Y[-1]) / np.dot(Y[-1], Y[-1])  # (7.20)
        q = gamma * q
    for a, rho, s, y in reversed(alphas):
        b = rho * np.dot(y, q)
        q = q + s * (a - b)
    return q


def _cubic_min(a, fa, dfa, b, fb, dfb):
    """Minimiser of the cubic through (a, fa, dfa), (b, fb, dfb) (N&W eq. 3.59); bisection if it is not usable."""
    d1 = dfa + dfb - 3.0 * (fa - fb) / (a - b)
    rad = d1 * d1 - dfa * dfb
    if not np.isfinite(rad) or rad < 0.0:
        return 0.5 * (a + b)
    d2 = np.sign(b - a) * np.sqrt(rad)
    den = dfb - dfa + 2.0 * d2
    if den == 0.0 or not np.isfinite(den):
        return 0.5 * (a + b)
    t = b - (b - a) * (dfb + d2 - d1) / den
    lo, hi = min(a, b), max(a, b)
    if not np.isfinite(t) or t <= lo + 0.1 * (hi - lo) or t >= hi - 0.1 * (hi - lo):
        return 0.5 * (a + b)
    return t


def wolfe_line_search(phi, f0, df0, step0=1.0, max_evals=MAX_LS_EVALS):
    """Algorithms 3.5 / 3.6: a step satisfying the strong Wolfe conditions for phi(t) -> (value, slope), phi'(0) = df0 < 0.
    Returns (t, value, slope, n_evals); falls back to the best point seen when the budget runs out."""
    evals = 0
    t_prev, f_prev, df_prev = 0.0, f0, df0
    t = step0
    best = (0.0, f0, df0)

    def zoom(lo, flo, dflo, hi, fhi, dfhi):
        nonlocal evals, best
        while evals < max_evals:
            t = _cubic_min(lo, flo, dflo, hi, fhi, dfhi)
            ft, dft = phi(t)
            evals += 1
            if np.isfinite(ft) and ft < best[1]:
                best = (t, ft, dft)
            if not np.isfinite(ft) or ft > f0 + C1 * t * df0 or ft >= flo:
                hi, fhi, dfhi = t, ft, dft
            else:
                if abs(dft) <= -C2 * df0:
                    return t, ft, dft
                if dft * (hi - lo) >= 0.0:
                    hi, fhi, dfhi = lo, flo, dflo
                lo, flo, dflo = t, ft, dft
        return best

    while evals < max_evals:
        ft, dft = phi(t)
        evals += 1
        if np.isfinite(ft) and ft < best[1]:
            best = (t, ft, dft)
        if not np.isfinite(ft) or ft > f0 + C1 * t * df0 or (evals > 1 and ft >= f_prev):
            r = zoom(t_prev, f_prev, df_prev, t, ft, dft)
            return r[0], r[1], r[2], evals
        if abs(dft) <= -C2 * df0:
            return t, ft, dft, evals
        if dft >= 0.0:
            r = zoom(t, ft, dft, t_prev, f_prev, df_prev)
            return r[0], r[1], r[2], evals
        t_prev, f_prev, df_prev = t, ft, dft
        t = 2.0 * t
    return best[0], best[1], best[2], evals


def lbfgs_minimize(value_and_grad, x0, max_iter=10, clip=1e6):
    """``max_iter`` L-BFGS steps on ``value_and_grad(x) -> (f, g)`` from x0, iterates clipped to +-clip after every step
    (optimizer.py:68).  Returns (x, trace of (x, f) per step)."""
    x = np.asarray(x0, dtype=np.float64).copy()
    f, g = value_and_grad(x)
    S, Y, trace = [], [], [(x.copy(), f)]
    for _ in range(max_iter):
        if not np.all(np.isfinite(g)) or np.linalg.norm(g) == 0.0:
            break
        p = -_two_loop(g, S, Y)
        df0 = float(np.dot(g, p))
        if not df0 < 0.0:  # not a descent direction (stale curvature pairs): restart from steepest descent
            S, Y = [], []
            p = -g
            df0 = float(np.dot(g, p))

        def phi(t, x=x, p=p):
            ft, gt = value_and_grad(x + t * p)
            return ft, float(np.dot(gt, p))

        t, ft, _, _ = wolfe_line_search(phi, f, df0)
        x_new = np.clip(x + t * p, -clip, clip)
        f_new, g_new = value_and_grad(x_new)
        s, y = x_new - x, g_new - g
        if np.dot(s, y) > 1e-12 * np.linalg.norm(s) * np.linalg.norm(y):  # curvature condition: keep the pair
            S.append(s)
            Y.append(y)
            if len(S) > MEMORY:
                S.pop(0)
                Y.pop(0)
        x, f, g = x_new, f_new, g_new
        trace.append((x.copy(), f))
    return x, trace


def optimize_suggestion(x0, fun_and_grad, max_iter=10, sign_quirk=False):
    """``_optimize_suggestion`` (optimizer.py:39-73) with this L-BFGS: ascend ``fun`` from x0 for ``max_iter`` steps.

    ``fun_and_grad(x) -> (fun(x), grad fun(x))``.  sign_quirk = False minimises ``-fun`` throughout (what the reference's
    comment says it does).  sign_quirk = True reproduces the reference's wiring (Q12): the FIRST value / gradient are those
    of ``-fun``, every later one comes from the line search, which was handed the un-negated ``fun`` - so from the second
    step on the iteration minimises ``+fun``."""
    def neg(x):
        v, g = fun_and_grad(x)
        return -v, -g

    if not sign_quirk:
        return lbfgs_minimize(neg, x0, max_iter)[0]
    x = np.asarray(x0, dtype=np.float64).copy()
    # step 1: direction from -fun (steepest descent on -fun = ascent direction), line search on +fun along it: the strong
    # Wolfe search needs a descent direction for the function it evaluates, which an ascent direction of fun is not - it
    # returns its best point, i.e. essentially stays put; later steps run L-BFGS on +fun
    v0, g0 = neg(x)
    p = -g0

    def phi(t):
        v, g = fun_and_grad(x + t * p)
        return v, float(np.dot(g, p))

    f_plus, g_plus = fun_and_grad(x)
    t, _, _, _ = wolfe_line_search(phi, f_plus, float(np.dot(g_plus, p))) if np.dot(g_plus, p) < 0 else (0.0, f_plus, 0.0, 0)
    x = np.clip(x + t * p, -1e6, 1e6)
    return lbfgs_minimize(fun_and_grad, x, max_iter - 1)[0]
