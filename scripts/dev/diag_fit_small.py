import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from bayex_b200 import engine
rng = np.random.default_rng(0)
for n in (24, 40, 48, 56, 64, 72, 80, 100, 128):
    d = 3
    X = rng.uniform(0, 1, (n, d)).astype(np.float32); Y = np.sin(3 * X.sum(1)).astype(np.float32)
    raw = np.concatenate([[-8.0], [0.0], np.zeros(d)]).astype(np.float32)
    post = engine.Posterior(X, Y, raw, want_tc=False)
    post.fit_hypers(steps=5); torch.cuda.synchronize()
    t0 = time.perf_counter(); post.fit_hypers(steps=200); torch.cuda.synchronize()
    print(n, "fused" if (n <= 80 and not os.environ.get("BX_NO_FIT_SMALL")) else "multi", round((time.perf_counter() - t0) / 200 * 1e3, 4), "ms/step", flush=True)
