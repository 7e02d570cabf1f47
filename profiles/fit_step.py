"""One NLML + gradient evaluation (= one step of posterior_fit without the AdamW update) at n = 4096 / 8192, inside an
NVTX range so `ncu --nvtx --nvtx-include "bx_fit/"` lists exactly its launches."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bayex_b200 import engine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
d = 8
rng = np.random.default_rng(0)
X = rng.uniform(0, 1, (n, d)).astype(np.float32)
Y = np.sin(3 * X.sum(1)).astype(np.float32)
raw = np.concatenate([[-8.0], [0.0], np.zeros(d)]).astype(np.float32)
post = engine.Posterior(X, Y, raw, want_tc=False)
post.nlml_grad(); torch.cuda.synchronize()
torch.cuda.nvtx.range_push("bx_fit")
out = post.nlml_grad()
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
print("nlml", out[0], "loss", out[1])
