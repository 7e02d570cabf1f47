"""Time bx_refine (30 trials) at n = 4096, d = 20 for S in argv (default 7 16 50) with CUDA events; dev tool."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import bayex_b200 as bayex
from bayex_b200 import engine as bengine
Ss = [int(a) for a in sys.argv[1:]] or [7, 16, 50]
d, n = 20, 4096
dom = {f"x{i}": bayex.domain.Real(0.0, 1.0) for i in range(d)}
dims = bengine.dims_array(dom)
X = bengine.threefry_fill(bayex.random.key(0), dims, d, 0, n).cpu().numpy()
Y = np.sin(3 * X.sum(1)).astype(np.float32)
raw = np.concatenate([[-8.0], [0.0], np.zeros(d)]).astype(np.float32)
post = bengine.Posterior(X, Y, raw).build()
for S in Ss:
    starts = bengine.threefry_fill(bayex.random.key(1), dims, d, 0, S)
    for _ in range(3):
        post.refine(starts.clone(), 30, "UCB", 0.01, 0.0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(post.stream)
    for _ in range(10):
        post.refine(starts.clone(), 30, "UCB", 0.01, 0.0)
    e1.record(post.stream)
    torch.cuda.synchronize()
    print(f"S={S}: {e0.elapsed_time(e1) / 10:.3f} ms per bx_refine (31 rounds)")
