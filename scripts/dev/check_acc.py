"""Accuracy and speed of the three scoring engines against the fp64 oracle: fp32 SIMT, the fast tensor mode alone (RAW) and
the accurate tensor mode (ACC: per-K-block chunks folded into TMEM-resident running sums).  sigma is recovered from UCB/LCB
with kappa = 1; error of sum v^2 = amp - sigma^2 in units of amp, over candidates next to and far from the observations."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import bayex_b200 as bayex
from bayex_b200 import _lib, engine
from oracle import gp as ogp, prng
from test_gpu_parity import make_problem

def sigma_of(post, key, dims, M, eng):
    u, _ = post.score_generated(key, dims, 0, M, "UCB", 1.0, 0.0, engine=eng)
    l, _ = post.score_generated(key, dims, 0, M, "LCB", 1.0, 0.0, engine=eng)
    u, l = u.cpu().numpy().astype(np.float64), l.cpu().numpy().astype(np.float64)
    return 0.5 * (u - l), 0.5 * (u + l)

for (n, d, nr, lo, hi) in [(700, 5, -4.0, -0.1, 1.1), (512, 6, -8.0, 0.0, 1.0), (2048, 10, -4.0, -0.1, 1.1), (2048, 10, -8.0, 0.0, 1.0),
                           (4096, 8, -8.0, 0.0, 1.0), (4096, 20, -8.0, 0.0, 1.0), (1000, 27, -4.0, 0.0, 1.0)]:
    X, Y, raw, P = make_problem(n, d, seed=31, noise_raw=nr)
    dom = {f"x{i}": bayex.domain.Real(lo, hi) for i in range(d)}
    dims = engine.dims_array(dom)
    post = engine.Posterior(X, Y, raw).build()
    amp = float(post.hyp()[0])
    key = prng.key(5)
    M = 8192
    cand = engine.threefry_fill(key, dims, d, 0, M).cpu().numpy()
    f64 = ogp.Factor(P, X, Y, np.ones(n), np.float64)
    f32 = ogp.Factor(P, X, Y, np.ones(n), np.float32)
    mu64, sd64 = f64.predict(cand.astype(np.float64))
    mu32, sd32 = f32.predict(cand)
    ssq64 = amp - sd64 ** 2
    print(f"n={n} d={d} noise_raw={nr}: var/amp median {np.median(sd64 ** 2) / amp:.3e} min {np.min(sd64 ** 2) / amp:.3e}; "
          f"fp32 NumPy oracle: max |d ssq|/amp {np.abs((amp - sd32.astype(np.float64) ** 2) - ssq64).max() / amp:.2e}, max |d mu| {np.abs(mu32 - mu64).max():.2e}")
    for name, eng in (("SIMT", _lib.ENGINE_SIMT), ("RAW ", _lib.ENGINE_TCGEN05_F16_RAW), ("ACC ", _lib.ENGINE_TCGEN05_F16_ACC)):
        sd, mu = sigma_of(post, key, dims, M, eng)
        e = ((amp - sd ** 2) - ssq64) / amp
        big = 1 << 18
        torch.cuda.synchronize(); t0 = time.perf_counter()
        post.score_generated(key, dims, 0, big, "UCB", 1.0, 0.0, engine=eng, keep_acq=False)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(f"   {name}: sum v^2 error / amp: mean {e.mean():+.2e} max|.| {np.abs(e).max():.2e} p99 {np.percentile(np.abs(e), 99):.2e}; "
              f"|d mu| max {np.abs(mu - mu64).max():.2e}; nan {int(np.isnan(sd).sum())}; 2^18 candidates in {1e3 * dt:.2f} ms")
