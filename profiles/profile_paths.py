"""Wall-clock of the non-headline paths: hyper-parameter fit per step, factor build, sample() at the small configs."""
import json, sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bayex_b200 as bayex
from bayex_b200 import engine

def sync(): torch.cuda.synchronize()
out = {}
rng = np.random.default_rng(0)
for n, d in [(23, 1), (64, 3), (512, 6), (1024, 8), (2048, 10), (4096, 20), (8192, 8)]:
    X = rng.uniform(0, 1, (n, d)).astype(np.float32)
    Y = np.sin(3 * X.sum(1)).astype(np.float32)
    raw = np.concatenate([[-8.0], [0.0], np.zeros(d)]).astype(np.float32)
    post = engine.Posterior(X, Y, raw)
    steps = 20 if n >= 4096 else 100
    post.fit_hypers(steps=3); sync()
    t0 = time.perf_counter(); post.fit_hypers(steps=steps); sync(); t_fit = (time.perf_counter() - t0) / steps
    t0 = time.perf_counter(); post.fit_hypers(steps=steps, use_graph=False); sync(); t_nog = (time.perf_counter() - t0) / steps
    post.build(); sync()
    t0 = time.perf_counter(); post.build(); sync(); t_build = time.perf_counter() - t0
    out[f"n{n}_d{d}"] = dict(fit_ms_per_step_graph=1e3 * t_fit, fit_ms_per_step_plain=1e3 * t_nog, factor_build_ms=1e3 * t_build,
                             fit_tflops=(n ** 3 + n * n * (3 * d + 4)) / t_fit / 1e12)
    print(n, d, out[f"n{n}_d{d}"], flush=True)
# sample() / fit() latency through the public API
for name, d, n0, M, acq in [("c1_readme", 1, 23, 10_000, "PI"), ("c2", 6, 512, 1 << 20, "EI"), ("c4", 10, 2048, 1 << 20, "PI")]:
    dom = {f"x{i}": bayex.domain.Real(0.0, 1.0) for i in range(d)}
    P0 = {k: rng.uniform(0, 1, n0).astype(np.float32) for k in dom}
    Y0 = sum(np.sin(5 * P0[k]) for k in dom).astype(np.float32)
    opt = bayex.Optimizer(dom, acq=acq, maximize=True)
    t0 = time.perf_counter(); st = opt.init(Y0, P0); sync(); t_init = time.perf_counter() - t0
    key = bayex.random.key(0)
    opt.sample(key, st, size=M); sync()
    ts = []
    for i in range(5):
        t0 = time.perf_counter(); new = opt.sample(bayex.random.fold_in(key, i), st, size=M); sync(); ts.append(time.perf_counter() - t0)
    t0 = time.perf_counter(); st2 = opt.fit(st, 0.5, new); sync(); t_fitcall = time.perf_counter() - t0
    out[name] = dict(init_s=t_init, sample_ms=1e3 * float(np.median(ts)), fit_call_ms=1e3 * t_fitcall, cand_per_s=M / float(np.median(ts)))
    print(name, out[name], flush=True)
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "gpurun_out", "profile_paths.json"), "w"), indent=1)
