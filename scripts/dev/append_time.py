"""Posterior.append (bx_factor_append, one bordered row) against Posterior.build (Gram + Cholesky + inverse) with CUDA events,
and Optimizer.fit(trainsteps=0 / 100) wall time, for n in argv (default 500 2000 4000 8100); dev tool."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import bayex_b200 as bayex
from bayex_b200 import engine as bengine
ns = [int(a) for a in sys.argv[1:]] or [400, 2000, 4000, 8100]
d, reps = 8, 20
dom = {f"x{i}": bayex.domain.Real(0.0, 1.0) for i in range(d)}
dims = bengine.dims_array(dom)
for n in ns:
    X = bengine.threefry_fill(bayex.random.key(0), dims, d, 0, n + reps + 3).cpu().numpy()
    Y = np.sin(3 * X.sum(1)).astype(np.float32)
    raw = np.concatenate([[-6.0], [0.0], np.zeros(d)]).astype(np.float32)
    post = bengine.Posterior(X[:n], Y[:n], raw)
    for _ in range(3):
        post.build()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(post.stream)
    for _ in range(5):
        post.build()
    e1.record(post.stream)
    torch.cuda.synchronize()
    t_build = e0.elapsed_time(e1) / 5
    for a in range(3):
        post.append(X[n + a], Y[n + a])
    torch.cuda.synchronize()
    e0.record(post.stream)
    for a in range(3, 3 + reps):
        post.append(X[n + a], Y[n + a])
    e1.record(post.stream)
    torch.cuda.synchronize()
    t_app = e0.elapsed_time(e1) / reps
    print(f"n={n} d={d}: build {t_build:.3f} ms, append {t_app:.3f} ms per observation ({t_build / t_app:.1f}x); "
          f"T is {post.np_ ** 2 * 4 / 1e6:.0f} MB", flush=True)
# the optimiser level: fit() with the reference's 100 AdamW steps against fit(trainsteps=0)
for n in [n_ for n_ in ns if n_ <= 4100]:
    opt = bayex.Optimizer(dom, acq="EI", maximize=True)
    X = bengine.threefry_fill(bayex.random.key(0), dims, d, 0, n).cpu().numpy()
    P0 = {k: X[:, j] for j, k in enumerate(dom)}
    st = opt.init(np.sin(3 * X.sum(1)).astype(np.float32), P0)
    new = opt.sample(bayex.random.key(1), st, size=4096)
    def rebuild():
        opt._cache.clear()
        opt.posterior(st)
    rebuild()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        rebuild()
    torch.cuda.synchronize()
    t_rebuild = (time.perf_counter() - t0) / 3
    res = {}
    for steps in (100, 0):
        opt.fit(st, 0.5, new, trainsteps=steps)  # warm
        rebuild()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            opt.fit(st, 0.5, new, trainsteps=steps)
            rebuild()  # fit(trainsteps=0) extended st's factor in place: give st its own factor back for the next repetition
        torch.cuda.synchronize()
        res[steps] = (time.perf_counter() - t0) / 3 - t_rebuild
    print(f"n={n}: Optimizer.fit wall time {res[100] * 1e3:.2f} ms with 100 AdamW steps, {res[0] * 1e3:.2f} ms with trainsteps=0 "
          f"(cold rebuild of the factor from a state: {t_rebuild * 1e3:.2f} ms)", flush=True)
