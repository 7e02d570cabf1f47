"""Generate ``tests/golden/reference_shim.npz`` by executing the reference's OWN source.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs ``/root/reference``;
the GPU box never runs this):

    python oracle/make_golden.py

``/root/reference/bayex`` is imported unchanged with ``oracle/_shim`` standing in
for the ``jax`` / ``optax`` namespaces (JAX itself is not installable here - see
``oracle/_shim/README.md``).  Sizes are tiny because the shim's ``vmap`` is a
Python loop.  Keys ending in ``_f64`` were produced with
``jax_enable_x64=True``; the rest in the reference's default float32.
"""

from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(HERE, "_shim"))
sys.path.insert(0, "/root/reference")

import jax  # noqa: E402  (the shim)
import jax.numpy as jnp  # noqa: E402
import bayex  # noqa: E402  (the unmodified reference)
from bayex import acq as racq, gp as rgp  # noqa: E402

OUT = {}


def put(name, value):
    OUT[name] = np.asarray(value)


def tag():
    return "_f64" if jnp.zeros(1).dtype == np.float64 else ""


def gp_fixtures():
    """tests/test_gp.py:57-62,83 and tests/test_acq.py:14-21 fixtures, with 0/1/5/10 padded slots."""
    t = tag()
    P = rgp.GPParams(0.1, 1.0, 1.0)
    x, xt = jnp.linspace(-5, 5, 10), jnp.linspace(-6, 6, 20)
    y = jnp.sin(x)
    for pad in (0, 1, 5, 10):
        xp = jnp.concatenate([x, jnp.zeros(pad)])
        yp = jnp.concatenate([y, jnp.zeros(pad)])
        mp = jnp.concatenate([jnp.ones(10), jnp.zeros(pad)])
        put(f"gp1d_nlml_pad{pad}{t}", rgp.marginal_likelihood(P, xp, yp, mp))
        mu, sd = rgp.predict(P, xp, yp, mp, xt=xt)
        put(f"gp1d_mean_pad{pad}{t}", mu)
        put(f"gp1d_std_pad{pad}{t}", sd)
    x, xt = jnp.linspace(-3, 3, 10), jnp.linspace(-4, 4, 5)
    y = jnp.sin(x)
    for pad in (0, 3):
        xp = jnp.concatenate([x, jnp.zeros(pad)])
        yp = jnp.concatenate([y, jnp.zeros(pad)])
        mp = jnp.concatenate([jnp.ones(10), jnp.zeros(pad)])
        for nm, fn in (("EI", racq.expected_improvement), ("PI", racq.probability_improvement),
                       ("UCB", racq.upper_confidence_bounds), ("LCB", racq.lower_confidence_bounds)):
            put(f"acq1d_{nm}_pad{pad}{t}", fn(xt, xp, yp, mp, P))


def multidim_case():
    """n_obs=24 padded to 30, d=3, array-shaped raw hypers; 40 test points (M x M product as written)."""
    t = tag()
    rng = np.random.default_rng(1234)
    n_obs, n, d, M = 24, 30, 3, 40
    X = np.zeros((n, d))
    X[:n_obs] = rng.uniform(-2, 2, (n_obs, d))
    Y = np.zeros(n)
    Y[:n_obs] = np.sin(X[:n_obs].sum(1)) + 0.1 * rng.standard_normal(n_obs)
    mask = np.arange(n) < n_obs
    XT = rng.uniform(-2.5, 2.5, (M, d))
    XT[:4] = X[:4] + 1e-3  # points close to observations: variance cancellation region
    P = rgp.GPParams(jnp.full((1, 1), -4.0), jnp.full((1, 1), 0.3), jnp.asarray([[0.2, -0.3, 0.7]]))
    if not t:
        for k, v in (("X", X), ("Y", Y), ("mask", mask), ("XT", XT)):
            put(f"md_{k}", v)
        put("md_params", np.array([-4.0, 0.3, 0.2, -0.3, 0.7]))
    Xj, Yj, XTj, mj = jnp.asarray(X), jnp.asarray(Y), jnp.asarray(XT), jnp.asarray(mask)
    put(f"md_nlml{t}", rgp.marginal_likelihood(P, Xj, Yj, mj))
    put(f"md_loss{t}", rgp.neg_log_likelihood(P, Xj, Yj, mj))
    mu, sd = rgp.predict(P, Xj, Yj, mj, xt=XTj)
    put(f"md_mean{t}", mu)
    put(f"md_std{t}", sd)
    for nm, fn in (("EI", racq.expected_improvement), ("PI", racq.probability_improvement),
                   ("UCB", racq.upper_confidence_bounds), ("LCB", racq.lower_confidence_bounds)):
        put(f"md_{nm}{t}", fn(XTj, Xj, Yj, mj, P))
    if t:  # finite-difference gradient of the reference's own objective (float64 only)
        g = jax.grad(rgp.neg_log_likelihood)(P, Xj, Yj, mj)
        put("md_lossgrad_fd_f64", np.concatenate([np.ravel(g.noise), np.ravel(g.amplitude), np.ravel(g.lengthscale)]))
        g = jax.grad(rgp.marginal_likelihood)(P, Xj, Yj, mj)
        put("md_nlmlgrad_fd_f64", np.concatenate([np.ravel(g.noise), np.ravel(g.amplitude), np.ravel(g.lengthscale)]))
        # the reference's own fit loop (gp.py:90-110) driven by FD gradients + the optax restatement
        Pf = rgp.posterior_fit(Yj, Xj, mj, P, trainsteps=20)
        put("md_fit20_f64", np.concatenate([np.ravel(Pf.noise), np.ravel(Pf.amplitude), np.ravel(Pf.lengthscale)]))


def optimizer_runs():
    """State machine + sample() of the reference (refinement = identity under the shim), float32."""
    # (a) README config: PI, Real(0,2), 3 init points, maximize (README.md:37-54)
    def f(x):
        return -(1.4 - 3 * x) * jnp.sin(18 * x)
    opt = bayex.Optimizer(domain={"x": bayex.domain.Real(0.0, 2.0)}, maximize=True, acq="PI")
    params = {"x": [0.0, 0.5, 1.0]}
    ys = [float(f(x)) for x in params["x"]]
    st = opt.init(ys, params)
    put("c1_init_gp", np.concatenate([np.ravel(v) for v in st.gp_params]))
    put("c1_init_ys", st.ys)
    put("c1_init_mask", st.mask)
    put("c1_init_x", st.params["x"])
    put("c1_init_best", np.array([st.best_score, float(st.best_params["x"])]))
    key = jax.random.key(42)
    props, gps, bests = [], [], []
    for step in range(9):  # crosses the 10 -> 20 buffer doubling (optimizer.py:224)
        key = jax.random.fold_in(key, step)
        new = opt.sample(key, st, size=96)
        y = f(**new)
        st = opt.fit(st, y, new)
        props.append(np.ravel(new["x"])[0])
        gps.append(np.concatenate([np.ravel(v) for v in st.gp_params]))
        bests.append(st.best_score)
    put("c1_props", np.array(props))
    put("c1_gps", np.array(gps))
    put("c1_bests", np.array(bests))
    put("c1_final_ys", st.ys)
    put("c1_final_mask", st.mask)
    put("c1_final_x", st.params["x"])

    # (b) mixed Real/Integer, minimise, each acquisition: one sample() on a fixed (unfitted) state
    dom = {"a": bayex.domain.Real(-5.0, 5.0), "k": bayex.domain.Integer(-10, 10), "b": bayex.domain.Real(0.0, 1.0)}
    rng = np.random.default_rng(7)
    n0 = 17
    P0 = {"a": rng.uniform(-5, 5, n0), "k": rng.integers(-10, 11, n0), "b": rng.uniform(0, 1, n0)}
    Y0 = (P0["a"] - 1.0) ** 2 + 0.3 * (P0["k"] - 2) ** 2 + P0["b"]
    for k, v in P0.items():
        put(f"c4_P_{k}", v)
    put("c4_Y", Y0)
    for acq in ("EI", "PI", "UCB", "LCB"):
        opt = bayex.Optimizer(domain=dom, maximize=False, acq=acq)
        st = opt.init(Y0, P0)
        put(f"c4_{acq}_gp", np.concatenate([np.ravel(v) for v in st.gp_params]))
        put(f"c4_{acq}_kdtype_is_int32", np.array(st.params["k"].dtype == np.int32))
        new = opt.sample(jax.random.key(5), st, size=80)
        put(f"c4_{acq}_prop", np.array([np.ravel(new[k])[0] for k in dom]))
        put(f"c4_{acq}_prop_dtypes_f32", np.array([new[k].dtype == np.float32 for k in dom]))
        st2 = opt.fit(st, 3.25, new)
        put(f"c4_{acq}_fit_k", st2.params["k"])
        put(f"c4_{acq}_fit_ys", st2.ys)
        put(f"c4_{acq}_fit_best", np.array([st2.best_score] + [float(st2.best_params[k]) for k in dom]))


def main():
    jax.config.update("jax_enable_x64", False)
    gp_fixtures()
    multidim_case()
    optimizer_runs()
    jax.config.update("jax_enable_x64", True)
    gp_fixtures()
    multidim_case()
    jax.config.update("jax_enable_x64", False)
    # the three closed-form KATs of tests/test_gp.py:10-12, evaluated by the reference function itself
    put("kat_expq", np.array([rgp.exp_quadratic(jnp.array([0.0]), jnp.array([0.0]), 1.0),
                              rgp.exp_quadratic(jnp.array([0.0]), jnp.array([1.0]), 1.0),
                              rgp.exp_quadratic(jnp.array([1.0]), jnp.array([1.0]), 0.0)]))
    path = os.path.join(ROOT, "tests", "golden", "reference_shim.npz")
    np.savez_compressed(path, **OUT)
    print(f"wrote {path}: {len(OUT)} arrays")


if __name__ == "__main__":
    main()
