import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import bayex_b200 as bayex
from bayex_b200 import _lib, engine
from oracle import acq as oacq, gp as ogp, prng
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
from test_gpu_parity import make_problem, RTOL
for (n, d, M, acq) in [(8192, 8, 700, "EI"), (4096, 8, 700, "EI"), (8192, 8, 700, "UCB")]:
    X, Y, raw, P = make_problem(n, d, seed=n * 3 + d)
    dom = {f"x{i}": bayex.domain.Real(-0.1, 1.1) for i in range(d)}
    dims = engine.dims_array(dom)
    post = engine.Posterior(X, Y, raw, want_tc=True).build()
    ystar = bayex.acq.ystar_for(acq, Y, np.ones(n)); param = bayex.acq.DEFAULT_PARAM[acq]
    key = prng.key(n)
    cand = engine.threefry_fill(key, dims, d, 5, M).cpu().numpy()
    f64, f32 = ogp.Factor(P, X, Y, np.ones(n), np.float64), ogp.Factor(P, X, Y, np.ones(n), np.float32)
    mu64, sd64 = f64.predict(cand.astype(np.float64)); mu32, sd32 = f32.predict(cand)
    a64 = oacq.acq_from_moments(acq, mu64, sd64, Y.astype(np.float64), np.ones(n))
    a32 = oacq.acq_from_moments(acq, mu32, sd32, Y, np.ones(n))
    print(f"n={n} d={d} {acq}: amp={post.hyp()[0]:.3f} noise={post.hyp()[1]:.2e} sd64 range [{sd64.min():.3e}, {sd64.max():.3e}] |a64| mean {np.abs(a64).mean():.3e}; fp32 oracle max|a32-a64| {np.abs(a32-a64).max():.3e}, sd rel err max {np.abs(sd32-sd64).max()/sd64.mean():.2e}")
    for name, eng in [("simt", _lib.ENGINE_SIMT), ("tf32", _lib.ENGINE_TCGEN05), ("tf32_pair", _lib.ENGINE_TCGEN05_PAIR), ("f16", _lib.ENGINE_TCGEN05_F16), ("f16_pair", _lib.ENGINE_TCGEN05_F16_PAIR), ("f16_shared", _lib.ENGINE_TCGEN05_F16_SHARED)]:
        v, _ = post.score_generated(key, dims, 5, M, acq, param, ystar, engine=eng)
        got = v.cpu().numpy().astype(np.float64)
        err = np.abs(got - a64)
        print(f"   {name:10s} max abs err {err.max():.3e}  rel-to-mean {err.max()/np.abs(a64).mean():.2e}  median {np.median(err):.2e}")
    mom = post.score_points(cand, acq=acq, param=param, ystar=ystar, want=("mu", "sigma"))
    print("   simt sigma max rel err", (np.abs(mom["sigma"].cpu().numpy() - sd64) / sd64).max(), "mu abs err", np.abs(mom["mu"].cpu().numpy() - mu64).max())
