"""Ablation of the accurate tensor mode (and the fast mode beside it): kernel time with the K* math and / or the MMA issue
switched off (BX_TC_ABLATE bit 0 / bit 1), n = 4096, 2^19 candidates, d = 8 and d = 20."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import bayex_b200 as bayex
from bayex_b200 import _lib, engine as bengine
M = 1 << 19
for d in (8, 20):
    n = 4096
    dom = {f"x{i}": bayex.domain.Real(0.0, 1.0) for i in range(d)}
    dims = bengine.dims_array(dom)
    X = bengine.threefry_fill(bayex.random.key(0), dims, d, 0, n).cpu().numpy()
    Y = np.sin(3 * X.sum(1)).astype(np.float32)
    raw = np.concatenate([[-8.0], [0.0], np.zeros(d)]).astype(np.float32)
    post = bengine.Posterior(X, Y, raw).build()
    key = bayex.random.key(1)
    for name, eng in (("fast", _lib.ENGINE_TCGEN05_F16_RAW), ("accurate", _lib.ENGINE_TCGEN05_F16_ACC)):
        row = []
        for ab in (0, 1, 2, 3):
            os.environ["BX_TC_ABLATE"] = str(ab)
            post.score_generated(key, dims, 0, M, "UCB", 0.01, 0.0, engine=eng, keep_acq=False)
            torch.cuda.synchronize(); t0 = time.perf_counter()
            post.score_generated(key, dims, 0, M, "UCB", 0.01, 0.0, engine=eng, keep_acq=False)
            torch.cuda.synchronize(); row.append(1e3 * (time.perf_counter() - t0))
        os.environ["BX_TC_ABLATE"] = "0"
        print(f"d={d} {name:9s}: full {row[0]:.2f} ms | no K* math {row[1]:.2f} | no MMA {row[2]:.2f} | neither {row[3]:.2f}   ({os.environ.get('BAYEX_B200_LIB', 'default lib')})")
