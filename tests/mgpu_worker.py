"""Worker of tests/test_multi_gpu.py (not a test module): one rank of a sharded sample() / fit sweep.

    torchrun --nproc-per-node N tests/mgpu_worker.py OUT_PREFIX        (N ranks, NCCL)
    python tests/mgpu_worker.py OUT_PREFIX                             (single process, no process group)

Every rank writes OUT_PREFIX.rank<r>.json: proposals (all four acquisitions, ragged candidate counts, fewer candidates than
ranks), the merged top-k indices, the refined points, a fit() with and without hyper-parameter steps and the winner of a
restart sweep."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    out_prefix = sys.argv[1]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import bayex_b200 as bayex

    rng = np.random.default_rng(0)
    n, d = 600, 5
    dom = {f"x{i}": (bayex.domain.Integer(-3, 3) if i == 2 else bayex.domain.Real(0.0, 1.0)) for i in range(d)}
    X = rng.uniform(0, 1, (n, d)).astype(np.float32)
    X[:, 2] = rng.integers(-3, 4, n)
    Y = (np.sin(3 * X.sum(1)) + 0.1 * rng.standard_normal(n)).astype(np.float32)
    gp = bayex.gp.GPParams(np.full((1, 1), -4.0, np.float32), np.full((1, 1), 0.3, np.float32),
                           np.array([[0.1, -0.2, 1.5, 0.0, 0.3]], np.float32))
    res = {"world": world, "rank": rank}
    for acq in ("EI", "PI", "UCB", "LCB"):
        opt = bayex.Optimizer(dom, acq=acq, maximize=True)
        st = bayex.optimizer.OptimizerState(params={f"x{i}": X[:, i].copy() for i in range(d)}, ys=Y, best_score=float(Y.max()),
                                            best_params={}, mask=np.ones(n, bool), gp_params=gp)
        for size in (200_003, 1000, 3, 1):
            new, info = opt.sample(bayex.random.key(7), st, size=size, details=True)
            res[f"{acq}_{size}"] = {"proposal": {k: float(v[0]) for k, v in new.items()}, "chosen": info["chosen"],
                                    "value": info["value"], "top_idxs": info["top_idxs"].tolist(),
                                    "optimized": np.asarray(info["optimized"], np.float64).round(12).tolist(),
                                    "opt_vals": np.asarray(info["opt_vals"], np.float64).tolist(),
                                    "best_rnd": list(info["best_rnd"])}
    # cold cache: the factor is rebuilt on rank 0 and broadcast
    opt = bayex.Optimizer(dom, acq="EI", maximize=True)
    st2 = opt.init(Y[:64], {f"x{i}": X[:64, i] for i in range(d)})
    st2 = opt.fit(st2, 0.5, {f"x{i}": np.array([X[100, i]], np.float32) for i in range(d)})
    res["after_fit"] = {k: float(v[0]) for k, v in opt.sample(bayex.random.key(3), st2, size=50_000).items()}
    res["after_fit_raw"] = [float(np.ravel(v)[j]) for v in st2.gp_params for j in range(np.size(v))]
    # hypers kept: every rank extends its own copy of the factor by one bordered row (no broadcast), same bits everywhere
    st3 = opt.fit(st2, 0.25, {f"x{i}": np.array([X[101, i]], np.float32) for i in range(d)}, trainsteps=0)
    assert opt.posterior(st3).n == 66
    res["after_append"] = {k: float(v[0]) for k, v in opt.sample(bayex.random.key(4), st3, size=50_000).items()}
    # BASELINE config 5 in miniature: restart sweep sharded over the ranks
    raws = np.concatenate([rng.uniform(-8, -2, (6, 1)), rng.uniform(-1, 1, (6, 1 + d))], axis=1).astype(np.float32)
    best, idx, table = bayex.gp.posterior_fit_restarts(Y[:300], X[:300], np.ones(300, bool), raws, trainsteps=20)
    res["sweep"] = {"best": int(idx), "table": np.asarray(table, np.float64).tolist()}
    json.dump(res, open(f"{out_prefix}.rank{rank}.json", "w"))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
