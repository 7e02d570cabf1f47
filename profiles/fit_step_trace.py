"""Per-kernel device times of NLML+gradient evaluations (posterior_fit steps without AdamW) from a PLAIN run (CUPTI via
torch.profiler: warm caches, real clocks, back-to-back launches) - complements the ncu launch list, whose per-launch
times are cold-cache and taken at whatever clock a short isolated kernel gets."""
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile

from bayex_b200 import engine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
d, reps = 8, 5
rng = np.random.default_rng(0)
X = rng.uniform(0, 1, (n, d)).astype(np.float32)
Y = np.sin(3 * X.sum(1)).astype(np.float32)
raw = np.concatenate([[-8.0], [0.0], np.zeros(d)]).astype(np.float32)
post = engine.Posterior(X, Y, raw, want_tc=False)
for _ in range(3):
    post.nlml_grad()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(reps):
        post.nlml_grad()
    torch.cuda.synchronize()
agg = collections.OrderedDict()
for ev in prof.events():
    if ev.device_type.name != "CUDA":
        continue
    name = ev.name.split("(")[0]
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
tot = sum(v[1] for v in agg.values())
print(f"# n={n} d={d}: device time per step {tot / reps / 1e3:.3f} ms (sum of kernels, {reps} steps)")
print("kernel,launches_per_step,us_per_step,share")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k},{c / reps:.1f},{t / reps:.1f},{t / tot:.4f}")
