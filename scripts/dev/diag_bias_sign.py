"""Signed relative error of sum v^2 from the tcgen05 fast mode ALONE (BX_HANDOVER_TAU=0 switches the hand-over to the
accurate mode off, so nothing is re-scored) against the fp64 oracle: the truncation bias of DESIGN.md section 3.  Dev tool."""
import os, sys
os.environ["BX_HANDOVER_TAU"] = "0"  # read once by the library at its first scoring call
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from bayex_b200 import _lib, engine
from oracle import gp as ogp
from test_gpu_parity import make_problem
for (n, d, nr) in [(700, 5, -4.0), (2048, 10, -4.0), (4096, 20, -4.0), (4096, 8, -8.0)]:
    X, Y, raw, P = make_problem(n, d, seed=31, noise_raw=nr)
    M = _lib.POINTS_TC_MIN + 777
    rng = np.random.default_rng(5)
    xt = rng.uniform(-0.1, 1.1, (M, d)).astype(np.float32)
    f64 = ogp.Factor(P, X, Y, np.ones(n), np.float64)
    mu64, sd64 = f64.predict(xt.astype(np.float64))
    post = engine.Posterior(X, Y, raw, want_tc=True).build()
    tc = post.score_points(xt, want=("mu", "sigma"))
    amp = post.hyp()[0]
    ss64 = amp - sd64 ** 2                         # sum v^2 (fp64)
    ss_tc = amp - tc["sigma"].cpu().numpy().astype(np.float64) ** 2
    rel = (ss_tc - ss64) / np.maximum(ss64, 1e-12)
    big = ss64 > 0.05 * amp
    print(f"n={n} d={d} noise_raw={nr}: sum v^2 signed rel err (tc - f64)/f64 over candidates with sum v^2 > 0.05 amp: mean {rel[big].mean():+.2e} median {np.median(rel[big]):+.2e} std {rel[big].std():.2e} p1 {np.percentile(rel[big],1):+.2e} p99 {np.percentile(rel[big],99):+.2e}  (#MMA per accumulator up to {3*n//16})")
