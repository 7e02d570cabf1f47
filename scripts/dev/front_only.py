"""One front half of sample() (score -> hand-over pass -> merge) at C3: target of `ncu --set full` on the scoring kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import bayex_b200 as bayex
from bayex_b200 import _lib, engine as bengine
M = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 24
eng = {"auto": _lib.ENGINE_AUTO, "acc": _lib.ENGINE_TCGEN05_F16_ACC, "raw": _lib.ENGINE_TCGEN05_F16_RAW}[sys.argv[2] if len(sys.argv) > 2 else "auto"]
d, n = 20, 4096
dom = {f"x{i}": bayex.domain.Real(0.0, 1.0) for i in range(d)}
dims = bengine.dims_array(dom)
X = bengine.threefry_fill(bayex.random.key(0), dims, d, 0, n).cpu().numpy()
x = X * 65.536 - 32.768
Y = (-20.0 * np.exp(-0.2 * np.sqrt(np.sum(x * x, axis=1) / d)) - np.exp(np.sum(np.cos(2 * np.pi * x), axis=1) / d) + 20.0 + np.e).astype(np.float32)
raw = np.concatenate([[-8.0], [0.0], np.zeros(d)]).astype(np.float32)
post = bengine.Posterior(X, -Y, raw)
post.fit_hypers(steps=100)
post.build()
for i in range(2):
    post.sample_front(bayex.random.key(1 + i), dims, 0, M, "UCB", 0.01, 0.0, 50, engine=eng)
torch.cuda.synchronize()
print("ok")
