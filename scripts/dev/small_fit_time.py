"""posterior_fit (100 AdamW steps) per-step time at small n, d in argv-free defaults; run with BX_NO_FIT_SWEEP=1 for the previous
paths (one-CTA shared-memory kernel up to n = 56, multi-kernel CUDA graph above); dev tool."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from bayex_b200 import engine
rng = np.random.default_rng(0)
for n, d in [(3, 1), (23, 1), (40, 3), (56, 4), (64, 3), (100, 6), (128, 8), (128, 20)]:
    X = rng.uniform(0, 1, (n, d)).astype(np.float32)
    Y = np.sin(3 * X.sum(1)).astype(np.float32)
    raw = np.concatenate([[-8.0], [0.0], np.zeros(d)]).astype(np.float32)
    post = engine.Posterior(X, Y, raw)
    post.fit_hypers(steps=5); torch.cuda.synchronize()
    post.set_raw(raw)
    t0 = time.perf_counter(); post.fit_hypers(steps=100); torch.cuda.synchronize(); t = (time.perf_counter() - t0) / 100
    r = post.raw_host()
    print(f"n={n} d={d}: {1e3 * t:.4f} ms per step; fitted raw[:3] {r[:3]}", flush=True)
