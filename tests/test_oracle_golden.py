"""The oracle restatement vs golden vectors made by executing the reference's own source
(``oracle/make_golden.py``), the reference's closed-form KATs and SURVEY section 4's fp64 values."""
import numpy as np
import pytest

from oracle import acq as oacq, gp as ogp, optimizer as oopt, domain as odom, prng

P1 = ogp.GPParams(0.1, 1.0, 1.0)


def _pad(x, y, pad):
    return (np.concatenate([x, np.zeros(pad, x.dtype)]), np.concatenate([y, np.zeros(pad, y.dtype)]),
            np.concatenate([np.ones(10), np.zeros(pad)]))


def test_exp_quadratic_kats(golden):
    # reference tests/test_gp.py:10-12
    assert np.isclose(ogp.exp_quadratic(np.array([0.0]), np.array([0.0]), 1.0), 1.0)
    assert np.isclose(ogp.exp_quadratic(np.array([0.0]), np.array([1.0]), 1.0), np.exp(-1.0))
    assert ogp.exp_quadratic(np.array([1.0]), np.array([1.0]), 0.0) == 0.0
    assert np.allclose(golden["kat_expq"], [1.0, np.exp(-1.0), 0.0])


def test_cov_shapes():
    # reference tests/test_gp.py:20-31
    x = np.linspace(0, 1, 10)
    assert ogp.cov(x, x, np.ones(10), np.ones(10)).shape == (10, 10)
    assert ogp.cov(x[:5], x, np.ones(5), np.ones(10)).shape == (5, 10)


@pytest.mark.parametrize("dt,tag,tol", [(np.float32, "", 2e-6), (np.float64, "_f64", 1e-12)])
@pytest.mark.parametrize("pad", [0, 1, 5, 10])
def test_gp1d_against_reference_source(golden, dt, tag, tol, pad):
    x = np.linspace(-5, 5, 10).astype(dt)
    xp, yp, mp = _pad(x, np.sin(x), pad)
    xt = np.linspace(-6, 6, 20).astype(dt)
    nl = ogp.marginal_likelihood(P1, xp, yp, mp, dtype=dt)
    assert np.isclose(nl, golden[f"gp1d_nlml_pad{pad}{tag}"], rtol=tol * 5, atol=0)
    mu, sd = ogp.predict(P1, xp, yp, mp, xt=xt, dtype=dt)
    assert np.allclose(mu, golden[f"gp1d_mean_pad{pad}{tag}"], rtol=0, atol=tol)
    assert np.allclose(sd, golden[f"gp1d_std_pad{pad}{tag}"], rtol=0, atol=tol)


def test_survey_fp64_values():
    # SURVEY section 4 table (independent fp64 restatement)
    x = np.linspace(-5, 5, 10)
    assert np.isclose(ogp.marginal_likelihood(P1, x, np.sin(x), np.ones(10), dtype=np.float64), 12.157260202570534, rtol=1e-10)
    mu, sd = ogp.predict(P1, x, np.sin(x), np.ones(10), xt=np.linspace(-6, 6, 20), dtype=np.float64)
    assert np.allclose(mu[:5], [0.31145065, 0.56587834, 0.69907323, 0.57937882, 0.23097205], atol=1e-8)
    assert np.allclose(sd[:5], [1.02032298, 0.77265735, 0.64636819, 0.64413596, 0.64634882], atol=1e-8)
    x, xt = np.linspace(-3, 3, 10), np.linspace(-4, 4, 5)
    want = {"EI": [0.08937889, 0.00012679, 0.00686294, 0.10004327, 0.08551781],
            "PI": [0.16512647, 0.0008496, 0.03225608, 0.29186109, 0.15940059],
            "UCB": [0.02210281, -0.70105025, 0.00543868, 0.71195999, -0.00169392],
            "LCB": [-2.61676726, -2.11167975, -1.40100486, -0.69866952, -2.64056398]}
    for nm, w in want.items():
        assert np.allclose(oacq.BY_NAME[nm](xt, x, np.sin(x), np.ones(10), P1, dtype=np.float64), w, atol=1e-8)
    assert np.isclose(ogp.softplus(0.1), 0.7443966600735709) and np.isclose(ogp.softplus(1.0), 1.3132616875182228)


@pytest.mark.parametrize("dt,tag,tol", [(np.float32, "", 3e-6), (np.float64, "_f64", 1e-12)])
@pytest.mark.parametrize("pad", [0, 3])
def test_acq1d_against_reference_source(golden, dt, tag, tol, pad):
    x = np.linspace(-3, 3, 10).astype(dt)
    xp, yp, mp = _pad(x, np.sin(x), pad)
    xt = np.linspace(-4, 4, 5).astype(dt)
    for nm, fn in oacq.BY_NAME.items():
        got = fn(xt, xp, yp, mp, P1, dtype=dt)
        assert got.shape == xt.shape and np.all(np.isfinite(got))  # reference tests/test_acq.py:36-47
        assert np.allclose(got, golden[f"acq1d_{nm}_pad{pad}{tag}"], rtol=0, atol=tol), nm


def _md(golden, dt):
    p = golden["md_params"]
    P = ogp.GPParams(np.full((1, 1), p[0]), np.full((1, 1), p[1]), p[2:].reshape(1, -1))
    return P, golden["md_X"].astype(dt), golden["md_Y"].astype(dt), golden["md_mask"], golden["md_XT"].astype(dt)


@pytest.mark.parametrize("dt,tag,tol", [(np.float32, "", 2e-5), (np.float64, "_f64", 1e-10)])
def test_multidim_against_reference_source(golden, dt, tag, tol):
    P, X, Y, mask, XT = _md(golden, dt)
    assert np.isclose(ogp.marginal_likelihood(P, X, Y, mask, dtype=dt), golden[f"md_nlml{tag}"], rtol=tol)
    assert np.isclose(ogp.neg_log_likelihood(P, X, Y, mask, dtype=dt), golden[f"md_loss{tag}"], rtol=tol)
    mu, sd = ogp.predict(P, X, Y, mask, xt=XT, dtype=dt)
    assert np.allclose(mu, golden[f"md_mean{tag}"], rtol=tol, atol=tol)
    # near-observation points sit in the fp32 cancellation region of amp - sum v^2: compare variances
    assert np.allclose(sd ** 2, golden[f"md_std{tag}"] ** 2, rtol=tol, atol=20 * tol)
    for nm, fn in oacq.BY_NAME.items():
        got = fn(XT, X, Y, mask, P, dtype=dt)
        assert np.allclose(got, golden[f"md_{nm}{tag}"], rtol=50 * tol, atol=50 * tol), nm
    # the diagonal-only posterior equals the reference's M x M formulation (quirk Q8)
    mu2, sd2 = ogp.predict_full_cov(P, X, Y, mask, XT, dtype=dt)
    assert np.allclose(mu, mu2, atol=tol) and np.allclose(sd ** 2, sd2 ** 2, atol=20 * tol)


def test_analytic_gradient_vs_reference_objective(golden):
    """Analytic gradient == finite differences of the reference's own neg_log_likelihood / marginal_likelihood."""
    P, X, Y, mask, _ = _md(golden, np.float64)
    g = ogp.loss_grad(P, X, Y, mask, dtype=np.float64)
    got = np.concatenate([[g.noise], [g.amplitude], g.lengthscale])
    assert np.allclose(got, golden["md_lossgrad_fd_f64"], rtol=2e-6, atol=1e-7)
    g = ogp.nlml_grad(P, X, Y, mask, dtype=np.float64)
    got = np.concatenate([[g.noise], [g.amplitude], g.lengthscale])
    assert np.allclose(got, golden["md_nlmlgrad_fd_f64"], rtol=2e-6, atol=1e-7)
    # fp32 analytic gradient stays within 1e-3 of it (conditioning of K^-1 in fp32)
    g32 = ogp.loss_grad(P, X.astype(np.float32), Y.astype(np.float32), mask, dtype=np.float32)
    got32 = np.concatenate([[g32.noise], [g32.amplitude], g32.lengthscale])
    assert np.allclose(got32, golden["md_lossgrad_fd_f64"], rtol=2e-3, atol=1e-3)


def test_fit_loop_vs_reference_loop(golden):
    """20 steps of the reference's posterior_fit (FD grads, optax restatement) vs the oracle loop (AdamW: unpinned)."""
    P, X, Y, mask, _ = _md(golden, np.float64)
    out = ogp.posterior_fit(Y, X, mask, P, trainsteps=20, dtype=np.float64)
    got = np.concatenate([np.ravel(out.noise), np.ravel(out.amplitude), np.ravel(out.lengthscale)])
    assert np.allclose(got, golden["md_fit20_f64"], rtol=0, atol=1e-7)


def _readme_f(x):
    return -(1.4 - 3 * x) * np.sin(18 * x)


def test_optimizer_state_machine_vs_reference_source(golden):
    """README config (PI, Real(0,2), maximise): init + 9 sample/fit steps incl. the 10->20 buffer doubling."""
    opt = oopt.Optimizer({"x": odom.Real(0.0, 2.0)}, acq="PI", maximize=True)
    params = {"x": [0.0, 0.5, 1.0]}
    ys = [float(_readme_f(np.float32(x))) for x in params["x"]]
    st = opt.init(ys, params)
    assert st.ys.shape == (10,) and st.mask.tolist() == [True] * 3 + [False] * 7
    assert np.allclose(st.ys, golden["c1_init_ys"], atol=1e-6)
    assert np.array_equal(st.mask, golden["c1_init_mask"])
    assert np.allclose(np.concatenate([np.ravel(v) for v in st.gp_params]), golden["c1_init_gp"], atol=2e-4)
    assert np.allclose([st.best_score, float(st.best_params["x"])], golden["c1_init_best"], atol=1e-6)
    key = prng.key(42)
    # Follow the reference trajectory: feed the golden proposals so one divergent argmax cannot cascade.
    for step in range(9):
        key = prng.fold_in(key, step)
        st_ref = st._replace(gp_params=ogp.GPParams(*np.split(golden["c1_gps"][step - 1], [1, 2]))) if step else st
        new = opt.sample(key, st_ref, size=96)
        assert new["x"].shape == (1,) and new["x"].dtype == np.float32
        assert np.isclose(new["x"][0], golden["c1_props"][step], atol=1e-6), step
        x = golden["c1_props"][step].astype(np.float32)
        st = opt.fit(st_ref, _readme_f(x), {"x": np.array([x])})
        # Adam divides by sqrt(v): where a gradient component crosses zero, fp32-vs-fp64 round-off in the
        # gradient is amplified to O(lr) per step, so fitted raw params agree to ~1e-2, not 1e-4.
        assert np.allclose(np.concatenate([np.ravel(v) for v in st.gp_params]), golden["c1_gps"][step], atol=1e-2)
        assert np.isclose(st.best_score, golden["c1_bests"][step], atol=1e-5)
    assert st.ys.shape == golden["c1_final_ys"].shape == (20,)
    assert np.array_equal(st.mask, golden["c1_final_mask"])
    assert np.allclose(st.params["x"], golden["c1_final_x"], atol=1e-6)
    assert np.allclose(st.ys, golden["c1_final_ys"], atol=1e-4)


@pytest.mark.parametrize("acq", ["EI", "PI", "UCB", "LCB"])
def test_mixed_domain_sample_fit_vs_reference_source(golden, acq):
    dom = {"a": odom.Real(-5.0, 5.0), "k": odom.Integer(-10, 10), "b": odom.Real(0.0, 1.0)}
    P0 = {k: golden[f"c4_P_{k}"] for k in dom}
    opt = oopt.Optimizer(dom, acq=acq, maximize=False)
    st = opt.init(golden["c4_Y"], P0)
    assert st.params["k"].dtype == np.int32 and bool(golden[f"c4_{acq}_kdtype_is_int32"])
    assert np.allclose(np.concatenate([np.ravel(v) for v in st.gp_params]), golden[f"c4_{acq}_gp"], atol=3e-4)
    gp_ref = golden[f"c4_{acq}_gp"]
    st = st._replace(gp_params=ogp.GPParams(gp_ref[:1].reshape(1, 1), gp_ref[1:2].reshape(1, 1), gp_ref[2:].reshape(1, -1)))
    new = opt.sample(prng.key(5), st, size=80)
    got = np.array([new[k][0] for k in dom])
    assert all(new[k].dtype == np.float32 for k in dom) and golden[f"c4_{acq}_prop_dtypes_f32"].all()
    assert got[1] == golden[f"c4_{acq}_prop"][1], "integer coordinate must match bit-exactly"
    assert np.allclose(got, golden[f"c4_{acq}_prop"], atol=1e-6)
    st2 = opt.fit(st, 3.25, {k: np.array([golden[f"c4_{acq}_prop"][i]], dtype=np.float32) for i, k in enumerate(dom)}, fit=False)
    assert st2.params["k"].dtype == np.float32  # quirk Q14: storage promoted by the first fit
    assert np.array_equal(st2.params["k"], golden[f"c4_{acq}_fit_k"])
    assert np.allclose(st2.ys, golden[f"c4_{acq}_fit_ys"], atol=1e-6)
    assert np.allclose([st2.best_score] + [float(st2.best_params[k]) for k in dom], golden[f"c4_{acq}_fit_best"], atol=1e-6)


def test_unknown_acq_raises():
    with pytest.raises(ValueError, match="Acquisition function magic is not implemented"):
        oopt.Optimizer({"x": odom.Real(-2, 2)}, acq="magic")


def test_acq_gradient_matches_finite_differences(golden):
    P, X, Y, mask, XT = _md(golden, np.float64)
    f = ogp.Factor(P, X, Y, mask, np.float64)
    pts = XT[4:10]
    for nm in ("EI", "PI", "UCB", "LCB"):
        val, grad = oopt.acq_value_and_grad(nm, f, Y, mask, pts)
        assert np.allclose(val, oacq.BY_NAME[nm](pts, X, Y, mask, P, dtype=np.float64, factor=f), rtol=1e-10, atol=1e-12)
        fd = np.zeros_like(grad)
        for k in range(pts.shape[1]):
            e = np.zeros(pts.shape[1]); e[k] = 1e-6
            fd[:, k] = (oacq.BY_NAME[nm](pts + e, X, Y, mask, P, dtype=np.float64, factor=f)
                        - oacq.BY_NAME[nm](pts - e, X, Y, mask, P, dtype=np.float64, factor=f)) / 2e-6
        assert np.allclose(grad, fd, rtol=1e-5, atol=1e-8), nm


def test_refine_ascent_is_monotone(golden):
    P, X, Y, mask, XT = _md(golden, np.float64)
    f = ogp.Factor(P, X, Y, mask, np.float64)
    for nm in ("EI", "UCB"):
        v0 = oacq.BY_NAME[nm](XT[:8], X, Y, mask, P, dtype=np.float64, factor=f)
        xb, fb = oopt.refine_ascent(nm, f, Y, mask, XT[:8])
        assert np.all(fb >= v0) and np.any(fb > v0)
        assert np.allclose(fb, oacq.BY_NAME[nm](xb, X, Y, mask, P, dtype=np.float64, factor=f), rtol=1e-9, atol=1e-12)


def test_adamw_restatement_matches_an_independent_implementation():
    """optax is unpinned and absent, so ``adamw_step`` (clip_by_global_norm(10) + adamw, gp.py:99) is restated from the
    published algorithm; torch.optim.AdamW + clip_grad_norm_ implement the same published algorithm independently
    (decoupled weight decay, bias-corrected moments, eps outside the square root).  60 steps on a gradient field that
    crosses the clipping threshold in both directions, fp64: the two trajectories agree to rounding."""
    import torch
    from oracle import gp as ogp
    rng = np.random.default_rng(3)
    P = 7
    A = rng.standard_normal((P, P))
    grad = lambda p, t: A @ p * (30.0 if t % 7 < 3 else 0.3) + np.sin(np.arange(P) + t)  # norms from ~1 to ~100
    p = rng.standard_normal(P)
    m, v, t = np.zeros(P), np.zeros(P), 0
    tp = torch.tensor(p, dtype=torch.float64, requires_grad=True)
    opt = torch.optim.AdamW([tp], lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-4)
    clipped = 0
    for step in range(60):
        g = grad(p, step)
        clipped += np.linalg.norm(g) > 10.0
        tp.grad = torch.tensor(grad(tp.detach().numpy(), step), dtype=torch.float64)
        torch.nn.utils.clip_grad_norm_([tp], 10.0)
        opt.step()
        p, m, v, t = ogp.adamw_step(p, g, m, v, t, 1e-3, np.float64)
        assert np.allclose(p, tp.detach().numpy(), rtol=0, atol=5e-9), (step, np.abs(p - tp.detach().numpy()).max())
    assert 0 < clipped < 60


def test_lbfgs_restatement_on_a_quadratic_and_in_the_reference_call_pattern():
    """oracle-only restatement of the refinement's optimiser (PARITY UNPINNED: optax's source is not
    available; this is Nocedal & Wright's L-BFGS + strong-Wolfe zoom).  Checks the algorithm itself (exact minimum of an
    ill-conditioned quadratic in <= d + 2 steps, Rosenbrock progress) and its use through the oracle's sample(): refined
    points never score below their starts, the proposal never below the best random candidate (the Q12 contract)."""
    from oracle import lbfgs as olbfgs, optimizer as oopt, domain as odom, prng
    rng = np.random.default_rng(0)
    d = 6
    A = rng.standard_normal((d, d))
    H = A @ A.T + np.diag(np.logspace(0, 3, d))
    b = rng.standard_normal(d)
    xs_true = np.linalg.solve(H, b)
    x, trace = olbfgs.lbfgs_minimize(lambda x: (0.5 * x @ H @ x - b @ x, H @ x - b), np.zeros(d), max_iter=40)
    assert np.allclose(x, xs_true, rtol=1e-6, atol=1e-8)
    assert all(t1[1] <= t0[1] + 1e-12 for t0, t1 in zip(trace[:-1], trace[1:])), "the Wolfe search must not increase f"

    def rosen(x):
        return (100.0 * (x[1] - x[0] ** 2) ** 2 + (1 - x[0]) ** 2,
                np.array([-400.0 * x[0] * (x[1] - x[0] ** 2) - 2 * (1 - x[0]), 200.0 * (x[1] - x[0] ** 2)]))
    x, trace = olbfgs.lbfgs_minimize(rosen, np.array([-1.2, 1.0]), max_iter=60)
    assert trace[-1][1] < 1e-10 and np.allclose(x, [1.0, 1.0], atol=1e-4)

    dom = {"a": odom.Real(0.0, 1.0), "b": odom.Real(0.0, 1.0), "c": odom.Real(0.0, 1.0)}
    P0 = {k: rng.uniform(0, 1, 30).astype(np.float32) for k in dom}
    Y0 = (np.sin(4 * P0["a"]) + P0["b"] ** 2 - P0["c"]).astype(np.float32)
    opt = oopt.Optimizer(dom, acq="EI", maximize=True)
    st = opt.init(Y0, P0, fit=False)
    key = prng.key(3)
    _, ident = opt.sample(key, st, size=2000, refine="identity", n_starts=8, details=True)
    _, lb = opt.sample(key, st, size=2000, refine="lbfgs", n_starts=8, details=True)
    assert np.all(lb["opt_vals"] >= ident["opt_vals"] - 1e-6 * np.abs(ident["opt_vals"]) - 1e-9)
    assert np.any(lb["opt_vals"] > ident["opt_vals"] * (1 + 1e-4)), "L-BFGS should improve at least one start"
    assert lb["all_vals"][lb["chosen"]] >= ident["acq_vals"].max() - 1e-9
    # the reference's wiring (mis-signed value_fn): finite, clipped, and the arg-max over optimised + random still
    # guarantees the proposal is at least the best random candidate - the contract the CUDA path is held to
    _, qk = opt.sample(key, st, size=2000, refine="lbfgs_quirk", n_starts=8, details=True)
    assert np.all(np.isfinite(qk["optimized"])) and np.all(np.abs(qk["optimized"]) <= 1e6)
    assert qk["all_vals"][qk["chosen"]] >= ident["acq_vals"].max() - 1e-9


def test_bordered_row_update_of_the_inverse_factor():
    """The identity behind bx_factor_append, checked in float64 on the oracle's own factor: appending an
    observation to gp.py:46-49's Gram matrix grows T = L^-1 by the row [-(l^T T) / lambda, 1 / lambda] with l = T k and
    lambda^2 = kappa - |l|^2, and leaves the rows above it untouched."""
    from scipy.linalg import solve_triangular
    rng = np.random.default_rng(3)
    n, d = 60, 3
    X = rng.uniform(0, 1, (n + 1, d))
    Y = np.sin(3 * X.sum(1))
    P = ogp.GPParams(-4.0, 0.3, np.array([0.1, -0.2, 0.4]))
    small = ogp.Factor(P, X[:n], Y[:n], np.ones(n), np.float64)
    full = ogp.Factor(P, X, Y, np.ones(n + 1), np.float64)
    T = solve_triangular(small.L, np.eye(n), lower=True)
    k = full.K[n, :n]
    l = T @ k
    lam = np.sqrt(full.K[n, n] - l @ l)
    row = np.concatenate([-(l @ T) / lam, [1.0 / lam]])
    T_full = solve_triangular(full.L, np.eye(n + 1), lower=True)
    assert np.allclose(T_full[:n, :n], T, rtol=0, atol=1e-12 * np.abs(T).max())
    assert np.allclose(T_full[n], row, rtol=1e-9, atol=1e-11 * np.abs(T_full).max())
    assert np.isclose(np.log(lam), np.sum(np.log(np.diag(full.L))) - np.sum(np.log(np.diag(small.L))), rtol=1e-10)
