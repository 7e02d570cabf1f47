"""Wall time per step of posterior_fit (100 captured AdamW steps, one restart) at the given sizes, d = 8.

    python profiles/fit_step_time.py 512 2048 4096 8192
    BX_CHOL_LOOKAHEAD_MIN=0 python profiles/fit_step_time.py ...   # without the look-ahead Cholesky
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bayex_b200 import engine

d = 8
for n in [int(a) for a in sys.argv[1:]] or [512, 2048, 4096, 8192]:
    rng = np.random.default_rng(0)
    X = rng.uniform(0, 1, (n, d)).astype(np.float32)
    Y = np.sin(3 * X.sum(1)).astype(np.float32)
    raw = np.concatenate([[-6.0], [0.0], np.zeros(d)]).astype(np.float32)
    post = engine.Posterior(X, Y, raw, want_tc=False)
    post.fit_hypers(steps=3); torch.cuda.synchronize()
    steps = 100 if n <= 4096 else 40
    best = 1e9
    for _ in range(2):
        post.set_raw(raw); torch.cuda.synchronize()
        t0 = time.perf_counter(); post.fit_hypers(steps=steps); torch.cuda.synchronize()
        best = min(best, (time.perf_counter() - t0) / steps)
    nl = post.nlml_grad()[0]
    print(f"n={n} ms_per_step={best * 1e3:.3f} tflops={n ** 3 / best / 1e12:.1f} nlml_after={nl:.6f} lookahead_min={os.environ.get('BX_CHOL_LOOKAHEAD_MIN', 'default')}", flush=True)
