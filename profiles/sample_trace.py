"""Where a sample() call goes at the small / mid configurations (C1 README, C2, C4): wall time vs summed device time per
kernel (CUPTI via torch.profiler, plain run)."""
import collections, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from torch.profiler import ProfilerActivity, profile
import bayex_b200 as bayex

rng = np.random.default_rng(0)
for name, d, n0, M, acq in [("c1_readme", 1, 23, 10_000, "PI"), ("c2", 6, 512, 1 << 20, "EI"), ("c4", 10, 2048, 1 << 20, "PI")]:
    dom = {f"x{i}": bayex.domain.Real(0.0, 1.0) for i in range(d)}
    P0 = {k: rng.uniform(0, 1, n0).astype(np.float32) for k in dom}
    Y0 = sum(np.sin(5 * P0[k]) for k in dom).astype(np.float32)
    opt = bayex.Optimizer(dom, acq=acq, maximize=True)
    st = opt.init(Y0, P0)
    key = bayex.random.key(0)
    for i in range(3):
        opt.sample(bayex.random.fold_in(key, i), st, size=M)
    torch.cuda.synchronize()
    reps = 5
    t0 = time.perf_counter()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for i in range(reps):
            opt.sample(bayex.random.fold_in(key, 10 + i), st, size=M)
        torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / reps
    agg = collections.OrderedDict()
    for ev in prof.events():
        if ev.device_type.name != "CUDA":
            continue
        a = agg.setdefault(ev.name.split("(")[0], [0, 0.0])
        a[0] += 1
        a[1] += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
    tot = sum(v[1] for v in agg.values())
    print(f"# {name}: n={n0} d={d} M={M}: wall {wall * 1e3:.3f} ms per sample() (under the profiler), device busy {tot / reps / 1e3:.3f} ms")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:8]:
        print(f"{k},{c / reps:.1f},{t / reps:.1f}us")
