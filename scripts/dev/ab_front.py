"""A/B of the front half of sample() on C3 (n = 4096, d = 20, 2^24 candidates, UCB): the tensor engine alone writing the
acquisition vector (round-1 behaviour) / alone with the fused top-k / the two-tier production path; plus the pieces after
it (gather, refinement, proposal) and C2 / C4 sample() latencies."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import bayex_b200 as bayex
from bayex_b200 import _lib, engine as bengine

def ev(fn, reps=3):
    fn(); torch.cuda.synchronize()
    out = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); out.append(1e3 * (time.perf_counter() - t0))
    return min(out)

d, n, M = 20, 4096, 1 << 24
dom = {f"x{i}": bayex.domain.Real(0.0, 1.0) for i in range(d)}
dims = bengine.dims_array(dom)
X = bengine.threefry_fill(bayex.random.key(0), dims, d, 0, n).cpu().numpy()
Y = np.sin(3 * X.sum(1)).astype(np.float32)
raw = np.concatenate([[-8.0], [0.0], np.zeros(d)]).astype(np.float32)
post = bengine.Posterior(X, Y, raw).build()
key = bayex.random.key(1)
print("C3 front half, ms:")
print("  raw engine, acquisition vector written, no top-k:", ev(lambda: post.score_generated(key, dims, 0, M, "UCB", 0.01, 0.0, keep_acq=True, engine=_lib.ENGINE_TCGEN05_F16_RAW)))
print("  raw engine, no vector (arg-max only):            ", ev(lambda: post.score_generated(key, dims, 0, M, "UCB", 0.01, 0.0, keep_acq=False, engine=_lib.ENGINE_TCGEN05_F16_RAW)))
print("  raw engine + fused top-k + merge:                ", ev(lambda: post.sample_front(key, dims, 0, M, "UCB", 0.01, 0.0, 50, engine=_lib.ENGINE_TCGEN05_F16_RAW)))
print("  production path (hand-over) + fused top-k + merge:", ev(lambda: post.sample_front(key, dims, 0, M, "UCB", 0.01, 0.0, 50, engine=_lib.ENGINE_AUTO)))
lst, tv, ti, _ = post.sample_front(key, dims, 0, M, "UCB", 0.01, 0.0, 50)
starts = post.gather(key, dims, ti)
print("  refinement of 50 starts, 30 rounds:", ev(lambda: post.refine(starts, 30, "UCB", 0.01, 0.0)))
print("  refinement of 7 starts, 30 rounds: ", ev(lambda: post.refine(starts[:7].contiguous(), 30, "UCB", 0.01, 0.0)))
for name, dd, nn, MM, acq in (("C2", 6, 512, 1 << 20, "EI"), ("C4-like real", 10, 2048, 1 << 20, "PI")):
    dm = {f"x{i}": bayex.domain.Real(0.0, 1.0) for i in range(dd)}
    Xc = bengine.threefry_fill(bayex.random.key(0), bengine.dims_array(dm), dd, 0, nn).cpu().numpy()
    Yc = (np.sin(3 * Xc.sum(1)) + 0.5 * np.cos(7 * Xc[:, 0])).astype(np.float32)
    opt = bayex.Optimizer(dm, acq=acq, maximize=True)
    st = opt.init(Yc, {f"x{i}": Xc[:, i] for i in range(dd)})
    for eng, nm in ((_lib.ENGINE_AUTO, "production"), (_lib.ENGINE_TCGEN05_F16_RAW, "tensor engine alone"), (_lib.ENGINE_SIMT, "SIMT alone")):
        opt.engine = eng
        print(f"{name} sample() [{nm}]: {ev(lambda: opt.sample(bayex.random.key(5), st, size=MM)):.2f} ms")
    p = opt.posterior(st)
    cnt = torch.zeros(1)
    print(f"{name} fit(): {ev(lambda: opt.fit(st, 0.1, {f'x{i}': np.array([0.5], np.float32) for i in range(dd)}), reps=2):.1f} ms")
