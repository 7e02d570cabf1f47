"""One bx_refine call (S starts from argv, default 7) at n = 4096, d = 20: target of the ncu launch list of the refinement."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import bayex_b200 as bayex
from bayex_b200 import engine as bengine
S = int(sys.argv[1]) if len(sys.argv) > 1 else 7
d, n = 20, 4096
dom = {f"x{i}": bayex.domain.Real(0.0, 1.0) for i in range(d)}
dims = bengine.dims_array(dom)
X = bengine.threefry_fill(bayex.random.key(0), dims, d, 0, n).cpu().numpy()
Y = np.sin(3 * X.sum(1)).astype(np.float32)
raw = np.concatenate([[-8.0], [0.0], np.zeros(d)]).astype(np.float32)
post = bengine.Posterior(X, Y, raw).build()
starts = bengine.threefry_fill(bayex.random.key(1), dims, d, 0, S)
torch.cuda.synchronize()
torch.cuda.nvtx.range_push("refine")
post.refine(starts, 2, "UCB", 0.01, 0.0)
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
print("ok")
