"""Small end-to-end pass over every kernel of the path (fit, factor build, fast + accurate tensor modes, SIMT engine, fused
top-k + merge, refinement, proposal, explicit-point scoring) - the target of `compute-sanitizer --tool memcheck`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import bayex_b200 as bayex
from bayex_b200 import _lib

rng = np.random.default_rng(0)
dom = {"a": bayex.domain.Real(0.0, 1.0), "k": bayex.domain.Integer(-3, 3), "b": bayex.domain.Real(-1.0, 1.0)}
n0 = 150
P0 = {"a": rng.uniform(0, 1, n0).astype(np.float32), "k": rng.integers(-3, 4, n0), "b": rng.uniform(-1, 1, n0).astype(np.float32)}
Y0 = (np.sin(4 * P0["a"]) + 0.1 * P0["k"] ** 2 + P0["b"]).astype(np.float32)
for acq in ("EI", "PI"):
    opt = bayex.Optimizer(dom, acq=acq)
    st = opt.init(Y0, P0)                                   # multi-kernel fit (n > 56), factor build
    for eng in (_lib.ENGINE_AUTO, _lib.ENGINE_SIMT, _lib.ENGINE_TCGEN05_F16_ACC):
        opt.engine = eng
        new = opt.sample(bayex.random.key(3), st, size=5000)  # fused top-k, second tier, merge, refinement, proposal
    st = opt.fit(st, 0.3, new)
xs = opt.param_space.to_array({k: st.params[k] for k in dom})
mu, sd = bayex.gp.predict(st.gp_params, xs, -st.ys, st.mask, rng.uniform(0, 1, (70000, 3)).astype(np.float32))  # two-tier explicit points
small = bayex.Optimizer({"x": bayex.domain.Real(0.0, 2.0)}, acq="UCB", maximize=True)
s2 = small.init([0.1, 0.5, 0.2], {"x": [0.0, 0.5, 1.0]})    # one-kernel fit (n <= 56)
small.sample(bayex.random.key(1), s2, size=300)
torch.cuda.synchronize()
print("sanity ok", float(mu[0]), float(sd[0]))
