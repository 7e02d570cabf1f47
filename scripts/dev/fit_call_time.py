"""Wall time of consecutive Optimizer.fit calls (C2-like: d = 6, EI) - first-call effects against steady state."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import bayex_b200 as bayex

for n0 in (508, 512, 2044):
    d = 6
    rng = np.random.default_rng(0)
    dom = {f"x{i}": bayex.domain.Real(0.0, 1.0) for i in range(d)}
    P0 = {k: rng.uniform(0, 1, n0).astype(np.float32) for k in dom}
    Y0 = sum(np.sin(5 * P0[k]) for k in dom).astype(np.float32)
    opt = bayex.Optimizer(domain=dom, maximize=True, acq="EI")
    t0 = time.perf_counter(); st = opt.init(Y0, P0); torch.cuda.synchronize(); t_init = time.perf_counter() - t0
    ts = []
    for i in range(6):
        new = {k: np.float32(rng.uniform()) for k in dom}
        t0 = time.perf_counter(); st = opt.fit(st, float(rng.uniform()), new); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    print(f"n0={n0} init {t_init * 1e3:.1f} ms; fit calls (ms):", " ".join(f"{t * 1e3:.1f}" for t in ts), flush=True)
