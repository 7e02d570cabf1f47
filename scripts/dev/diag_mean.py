"""Posterior mean of the SIMT engine against the fp64 oracle, in units of the test tolerance
max(1e-4 max(|mu|, mean|mu|), 4 |mu32 - mu64|), and the posterior build time."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from bayex_b200 import engine
from oracle import gp as ogp
from test_gpu_parity import make_problem
for (n, d, nr) in [(512, 6, -4.0), (2048, 10, -4.0), (2048, 10, -8.0), (4096, 8, -4.0)]:
    X, Y, raw, P = make_problem(n, d, seed=7 * n + d, noise_raw=nr)
    M = 2048
    cand = np.random.default_rng(n).uniform(-0.1, 1.1, (M, d)).astype(np.float32)
    f64, f32 = ogp.Factor(P, X, Y, np.ones(n), np.float64), ogp.Factor(P, X, Y, np.ones(n), np.float32)
    mu64, _ = f64.predict(cand.astype(np.float64)); mu32, _ = f32.predict(cand)
    post = engine.Posterior(X, Y, raw, want_tc=False).build(); torch.cuda.synchronize()
    t0 = time.perf_counter(); post.build(); torch.cuda.synchronize(); tb = time.perf_counter() - t0
    mu = post.score_points(cand, want=("mu",))["mu"].cpu().numpy().astype(np.float64)
    e = np.abs(mu - mu64); tol = np.maximum(1e-4 * np.maximum(np.abs(mu64), np.abs(mu64).mean()), 4 * np.abs(mu32 - mu64))
    print(f"n={n} d={d} noise_raw={nr}: max|err| {e.max():.2e} (fp32 oracle {np.abs(mu32 - mu64).max():.2e}); max err/tol {np.max(e / tol):.3f}; build {tb * 1e3:.2f} ms", flush=True)
