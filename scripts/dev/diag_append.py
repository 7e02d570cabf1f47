"""Appended factor against a factor rebuilt from scratch, both against the fp64 oracle (variance at explicit points); dev tool."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import numpy as np
from bayex_b200 import engine
from oracle import gp as ogp
from test_gpu_parity import make_problem

for n, d, nadd in [(37, 1, 3), (300, 3, 12), (120, 2, 8), (20, 2, 3), (120, 4, 8), (100, 3, 8), (500, 6, 12), (1000, 6, 5), (2040, 10, 8), (4000, 20, 16)]:
    N = n + nadd
    X, Y, raw, P = make_problem(N, d, seed=11 * n + d)
    rng = np.random.default_rng(n)
    X[n::2] = X[:len(X[n::2])] + rng.normal(0, 2e-3, X[n::2].shape).astype(np.float32)
    Y[n::2] = (np.sin(3 * X[n::2].sum(1)) + 0.1 * rng.standard_normal(len(X[n::2]))).astype(np.float32)
    post = engine.Posterior(X[:n], Y[:n], raw, want_tc=False).build()
    for a in range(n, N):
        post.append(X[a], Y[a])
    full = engine.Posterior(X, Y, raw, want_tc=False).build()
    mask = np.ones(N)
    f64, f32 = ogp.Factor(P, X, Y, mask, np.float64), ogp.Factor(P, X, Y, mask, np.float32)
    M = 3000
    cand = rng.uniform(-0.1, 1.1, (M, d)).astype(np.float32)
    cand[:nadd] = X[n:] + 1e-3
    cand[nadd:2 * nadd] = X[n:]
    mu64, sd64, var64 = f64.predict(cand.astype(np.float64), return_var=True)
    mu32, sd32, var32 = f32.predict(cand, return_var=True)
    r64, r32 = np.maximum(var64, 1e-10), np.maximum(var32, 1e-10)
    tol = np.maximum(1e-4 * np.maximum(r64, r64.mean()), 4 * np.abs(r32.astype(np.float64) - r64))
    line = f"n={n} d={d} nadd={nadd}:"
    for name, p in (("appended", post), ("rebuilt", full)):
        out = p.score_points(cand, acq="EI", param=0.01, ystar=float(Y.max()))
        var = out["sigma"].cpu().numpy().astype(np.float64) ** 2
        err = np.abs(var - r64)
        bad = err > tol
        line += f"  {name}: {bad.sum()} outside, worst err/tol {np.max(err / tol):.2f} (first 2*nadd rows: {bad[:2 * nadd].sum()})"
    dT = (post._bufs["T"] - full._bufs["T"]).abs().max().item() / full._bufs["T"].abs().max().item()
    print(line + f"  max|T_app - T_full|/max|T| {dT:.2e}", flush=True)
