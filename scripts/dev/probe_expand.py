"""Probe: C3 front half time + fast-mode-alone accuracy against the SIMT engine (run once per library variant)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import bayex_b200 as bayex
from bayex_b200 import _lib, engine as bengine
d, n, M = 20, 4096, 1 << 24
dom = {f"x{i}": bayex.domain.Real(0.0, 1.0) for i in range(d)}
dims = bengine.dims_array(dom)
X = bengine.threefry_fill(bayex.random.key(0), dims, d, 0, n).cpu().numpy()
Y = np.sin(3 * X.sum(1)).astype(np.float32)
raw = np.concatenate([[-8.0], [0.0], np.zeros(d)]).astype(np.float32)
post = bengine.Posterior(X, Y, raw).build()
key = bayex.random.key(1)
def ev(fn, reps=3):
    fn(); torch.cuda.synchronize(); out = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); out.append(1e3 * (time.perf_counter() - t0))
    return min(out)
print("lib:", os.environ.get("BAYEX_B200_LIB", "default"))
print("  production front half 2^24: %.1f ms" % ev(lambda: post.sample_front(key, dims, 0, M, "UCB", 0.01, 0.0, 50)))
m = 1 << 20
a, _ = post.score_generated(key, dims, 0, m, "LCB", 2.576, 0.0, keep_acq=True, engine=_lib.ENGINE_TCGEN05_F16_RAW)
b, _ = post.score_generated(key, dims, 0, m, "LCB", 2.576, 0.0, keep_acq=True, engine=_lib.ENGINE_SIMT)
a = a.cpu().numpy().astype(np.float64); b = b.cpu().numpy().astype(np.float64)
rel = np.abs(a - b) / np.maximum(np.abs(b), 1e-30)
print("  fast mode alone vs SIMT, LCB(2.576) over 2^20: max rel %.3e  p99.9 %.3e  median %.3e" % (rel.max(), np.percentile(rel, 99.9), np.median(rel)))
print("  raw front half 2^24 (fast mode alone): %.1f ms" % ev(lambda: post.sample_front(key, dims, 0, M, "UCB", 0.01, 0.0, 50, engine=_lib.ENGINE_TCGEN05_F16_RAW)))
