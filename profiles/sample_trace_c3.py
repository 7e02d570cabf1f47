"""Where one sample() goes at C3 (n = 4096, d = 20, UCB) for a shard of M candidates (default 2^21 = one rank of an 8-GPU
run): host-side phases (wall clock with a sync after each phase - NOT how sample() runs, it never syncs in between) and the
per-kernel device time of a plain call (CUPTI via torch.profiler)."""
import collections, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from torch.profiler import ProfilerActivity, profile
import bayex_b200 as bayex
from bayex_b200 import engine as bengine, acq as boacq

M = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 21
d, n = 20, 4096
dom = {f"x{i}": bayex.domain.Real(0.0, 1.0) for i in range(d)}
dims = bengine.dims_array(dom)
X = bengine.threefry_fill(bayex.random.key(0), dims, d, 0, n).cpu().numpy()
Y = np.sin(3 * X.sum(1)).astype(np.float32)
gp = bayex.gp.GPParams(np.full((1, 1), -8.0, np.float32), np.zeros((1, 1), np.float32), np.zeros((1, d), np.float32))
opt = bayex.Optimizer(dom, acq="UCB", maximize=False)
st = bayex.optimizer.OptimizerState(params={f"x{i}": X[:, i].copy() for i in range(d)}, ys=Y, best_score=float(Y.min()), best_params={},
                                    mask=np.ones(n, bool), gp_params=gp)
key = bayex.random.key(1)
for i in range(3):
    opt.sample(bayex.random.fold_in(key, i), st, size=M)
torch.cuda.synchronize()

def wall(fn, reps=5):
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); ts.append(1e3 * (time.perf_counter() - t0))
    return min(ts), r

print(f"# C3 shard of {M} candidates")
t, _ = wall(lambda: opt.sample(bayex.random.fold_in(key, 7), st, size=M))
print(f"sample() wall: {t:.3f} ms")
t, post = wall(lambda: opt.posterior(st)); print(f"  posterior(state) lookup (digest of the state arrays): {t:.3f} ms")
t, ystar = wall(lambda: boacq.ystar_for("UCB", -np.asarray(st.ys, np.float32), st.mask)); print(f"  ystar_for: {t:.3f} ms")
t, fr = wall(lambda: post.sample_front(key, dims, 0, M, "UCB", 0.01, 0.0, 50)); print(f"  sample_front (score + hand-over + merge): {t:.3f} ms")
lst, tv, ti, _ = fr
t, starts = wall(lambda: post.gather(key, dims, ti)); print(f"  gather 50 starts: {t:.3f} ms")
t, _ = wall(lambda: post.refine(starts, 30, "UCB", 0.01, 0.0)); print(f"  refine 50 starts x 30 rounds: {t:.3f} ms")
t, _ = wall(lambda: post.refine(starts[:7].contiguous(), 30, "UCB", 0.01, 0.0)); print(f"  refine 7 starts x 30 rounds: {t:.3f} ms")
x, v = post.refine(starts, 30, "UCB", 0.01, 0.0)
t, out = wall(lambda: post.propose(key, dims, x, v, 50, lst[50:])); print(f"  propose: {t:.3f} ms")
t, _ = wall(lambda: out.cpu().numpy()); print(f"  read-back: {t:.3f} ms")

reps = 3
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for i in range(reps):
        opt.sample(bayex.random.fold_in(key, 10 + i), st, size=M)
    torch.cuda.synchronize()
agg = collections.OrderedDict()
for ev in prof.events():
    if ev.device_type.name != "CUDA":
        continue
    a = agg.setdefault(ev.name.split("(")[0][:70], [0, 0.0])
    a[0] += 1
    a[1] += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
tot = sum(v[1] for v in agg.values())
print(f"# device busy {tot / reps / 1e3:.3f} ms per sample(); kernel, launches per call, device time per call")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
    print(f"{k},{c / reps:.1f},{t / reps:.1f}us")
