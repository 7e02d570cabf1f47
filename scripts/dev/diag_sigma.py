import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import bayex_b200 as bayex
from bayex_b200 import _lib, engine
from oracle import gp as ogp
from test_gpu_parity import make_problem
for (n, d) in [(700, 5), (4096, 20), (2048, 10)]:
    X, Y, raw, P = make_problem(n, d, seed=31)
    M = _lib.POINTS_TC_MIN + 777
    rng = np.random.default_rng(5)
    xt = rng.uniform(-0.1, 1.1, (M, d)).astype(np.float32); xt[:4] = X[:4] + 1e-3
    f64, f32 = ogp.Factor(P, X, Y, np.ones(n), np.float64), ogp.Factor(P, X, Y, np.ones(n), np.float32)
    mu64, sd64 = f64.predict(xt.astype(np.float64)); mu32, sd32 = f32.predict(xt)
    post = engine.Posterior(X, Y, raw, want_tc=True).build()
    tc = post.score_points(xt, want=("mu", "sigma"))                       # >= POINTS_TC_MIN -> tcgen05
    k = _lib.POINTS_TC_MIN - 1
    sm = post.score_points(xt[:k], want=("mu", "sigma"))                   # SIMT
    amp = post.hyp()[0]
    def stats(name, sd, mu, sl):
        rs = np.abs(sd - sd64[sl]) / sd64[sl]; am = np.abs(mu - mu64[sl]) / np.maximum(np.abs(mu64[sl]), np.abs(mu64).mean())
        print(f"  {name:12s} sigma rel err: median {np.median(rs):.1e} p99 {np.percentile(rs,99):.1e} max {rs.max():.1e} (at sd={sd64[sl][np.argmax(rs)]:.3f}) | var/amp abs err max {np.abs(sd.astype(np.float64)**2 - sd64[sl]**2).max()/amp:.1e} | mu rel err p99 {np.percentile(am,99):.1e} max {am.max():.1e}")
    print(f"n={n} d={d} amp={amp:.3f} sd64 in [{sd64.min():.3f},{sd64.max():.3f}]")
    stats("oracle fp32", sd32, mu32, slice(None)); stats("tcgen05", tc["sigma"].cpu().numpy(), tc["mu"].cpu().numpy(), slice(None)); stats("simt", sm["sigma"].cpu().numpy(), sm["mu"].cpu().numpy(), slice(0, k))
