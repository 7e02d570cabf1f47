"""Kernel time of the pair and the transposed engine with the K* math and / or the MMA issue switched off (BX_TC_ABLATE:
1 = no K* math, 2 = no MMA, 3 = neither) at n = 4096, d = 20, 2^21 candidates."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import bayex_b200 as bayex
from bayex_b200 import _lib, engine
from test_gpu_parity import make_problem

key = bayex.random.key(3)
n, d, M = 4096, 20, 1 << 21
X, Y, raw, P = make_problem(n, d, seed=1)
dims = engine.dims_array({f"x{i}": bayex.domain.Real(0.0, 1.0) for i in range(d)})
post = engine.Posterior(X, Y, raw).build()
for name, eng in (("pair", _lib.ENGINE_TCGEN05_F16_PAIR), ("transposed", _lib.ENGINE_TCGEN05_F16_T)):
    for ab in (0, 1, 2, 3):
        os.environ["BX_TC_ABLATE"] = str(ab)
        post.score_generated(key, dims, 0, M, "UCB", 0.01, 0.0, engine=eng); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(2):
            post.score_generated(key, dims, 0, M, "UCB", 0.01, 0.0, engine=eng)
        torch.cuda.synchronize()
        print(f"{name} ablate={ab}: {(time.perf_counter() - t0) / 2 * 1e3:.1f} ms", flush=True)
