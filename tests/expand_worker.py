"""Worker of test_expanded_distance_form (not a test module): the fast mode of the fp16 tcgen05 engine ALONE against the fp32
SIMT engine on the candidates the production path lets it finalise (variance >= amp / 2), for the squared-distance form the
environment selects (BX_EXPAND_SMAX is read once per process: 0 = direct form, large = expanded form whatever the norms).

    python tests/expand_worker.py OUT.json
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import bayex_b200 as bayex
    from bayex_b200 import _lib, engine as bengine
    from test_gpu_parity import make_problem

    out = {}
    for n, d in ((512, 6), (1300, 20), (700, 27), (384, 32)):
        X, Y, raw, _ = make_problem(n, d, seed=100 + d)
        dom = {f"x{i}": bayex.domain.Real(-0.2, 1.2) for i in range(d)}
        dims = bengine.dims_array(dom)
        post = bengine.Posterior(X, Y, raw, want_tc=True).build()
        key = bayex.random.key(9)
        m = 1 << 16
        amp = float(post.hyp()[0])
        var = {}
        for name, eng in (("fast", _lib.ENGINE_TCGEN05_F16_RAW), ("simt", _lib.ENGINE_SIMT)):
            mu, _ = post.score_generated(key, dims, 0, m, "UCB", 0.0, 0.0, keep_acq=True, engine=eng)
            ms, _ = post.score_generated(key, dims, 0, m, "UCB", 1.0, 0.0, keep_acq=True, engine=eng)
            sd = (ms - mu).cpu().numpy().astype(np.float64)
            var[name] = sd * sd
            out[f"{n}x{d}:{name}:mu_checksum"] = float(np.abs(mu.cpu().numpy().astype(np.float64)).sum())
        tier = var["simt"] >= 0.5 * amp
        rel = np.abs(var["fast"] - var["simt"])[tier] / var["simt"][tier]
        out[f"{n}x{d}:fast_tier_fraction"] = float(tier.mean())
        out[f"{n}x{d}:max_rel_var_err"] = float(rel.max()) if tier.any() else 0.0
    json.dump(out, open(sys.argv[1], "w"))


if __name__ == "__main__":
    main()
