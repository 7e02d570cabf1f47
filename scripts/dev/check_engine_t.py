"""Experimental transposed-tiling engine (BX_ENGINE_TCGEN05_F16_T, score_tc16t.cu) against the production pair engine:
values, arg-max, then kernel time at a C3-like size.  Run under `timeout`: a protocol bug traps after a bounded wait."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import bayex_b200 as bayex
from bayex_b200 import _lib, engine
from test_gpu_parity import make_problem

key = bayex.random.key(3)
for (n, d, M, acq) in [(600, 5, 20000, "UCB"), (2048, 10, 50001, "EI"), (1100, 20, 9000, "LCB")]:
    X, Y, raw, P = make_problem(n, d, seed=5 * n + d)
    dom = {f"x{i}": bayex.domain.Real(0.0, 1.0) for i in range(d)}
    dims = engine.dims_array(dom)
    post = engine.Posterior(X, Y, raw).build()
    ystar = bayex.acq.ystar_for(acq, Y, np.ones(n)); param = bayex.acq.DEFAULT_PARAM[acq]
    vp, bp = post.score_generated(key, dims, 7, M, acq, param, ystar, engine=_lib.ENGINE_TCGEN05_F16_PAIR)
    vt, bt = post.score_generated(key, dims, 7, M, acq, param, ystar, engine=_lib.ENGINE_TCGEN05_F16_T)
    torch.cuda.synchronize()
    p, t = vp.cpu().numpy().astype(np.float64), vt.cpu().numpy().astype(np.float64)
    print(f"n={n} d={d} M={M} {acq}: finite {np.isfinite(t).all()} max|t-p| {np.abs(t - p).max():.3e} (scale {np.abs(p).mean():.3e}) "
          f"argmax equal {int(bp.item()) == int(bt.item())}", flush=True)
n, d, M = 4096, 20, 1 << 21
X, Y, raw, P = make_problem(n, d, seed=1)
dom = {f"x{i}": bayex.domain.Real(0.0, 1.0) for i in range(d)}
dims = engine.dims_array(dom)
post = engine.Posterior(X, Y, raw).build()
for name, eng in (("pair", _lib.ENGINE_TCGEN05_F16_PAIR), ("transposed", _lib.ENGINE_TCGEN05_F16_T)):
    post.score_generated(key, dims, 0, M, "UCB", 0.01, 0.0, engine=eng); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        post.score_generated(key, dims, 0, M, "UCB", 0.01, 0.0, engine=eng)
    torch.cuda.synchronize()
    print(f"{name}: {(time.perf_counter() - t0) / 3 * 1e3:.1f} ms per 2^21 candidates at n=4096 d=20", flush=True)
