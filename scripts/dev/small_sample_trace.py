"""Per-phase device timeline (BX_TRACE_SAMPLE) and wall time of sample() in the README regime (n <= 256); dev tool."""
import os, sys, time
os.environ["BX_TRACE_SAMPLE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import bayex_b200 as bayex
for n0, d, M, acq in [(23, 1, 10_000, "PI"), (100, 3, 10_000, "EI"), (120, 6, 100_000, "EI"), (200, 4, 10_000, "EI"), (250, 6, 100_000, "UCB")]:
    rng = np.random.default_rng(0)
    dom = {f"x{i}": bayex.domain.Real(0.0, 1.0) for i in range(d)}
    P0 = {k: rng.uniform(0, 1, n0).astype(np.float32) for k in dom}
    Y0 = sum(np.sin(5 * P0[k]) for k in dom).astype(np.float32)
    opt = bayex.Optimizer(dom, acq=acq, maximize=True)
    st = opt.init(Y0, P0)
    key = bayex.random.key(0)
    print(f"== n={n0} d={d} M={M} {acq}", flush=True)
    for i in range(3):
        opt.sample(bayex.random.fold_in(key, i), st, size=M)
    torch.cuda.synchronize()
    os.environ.pop("BX_TRACE_SAMPLE")
    ts = []
    for i in range(20):
        t0 = time.perf_counter(); new = opt.sample(bayex.random.fold_in(key, 10 + i), st, size=M); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    print(f"   sample() wall: median {1e3 * np.median(ts):.3f} ms, min {1e3 * min(ts):.3f} ms; proposal {[float(v[0]) for v in new.values()]}", flush=True)
    os.environ["BX_TRACE_SAMPLE"] = "1"
