"""BASELINE config 5: GP hyper-parameter fit sweep - RBF marginal likelihood + Cholesky at n = 1024 / 4096 / 8192, d = 8,
32 random restarts x 100 AdamW steps, restarts sharded over the ranks (torchrun) - timing of the whole sweep.

    python profiles/c5_fit_sweep.py [n ...]                      # one GPU
    BX_C5_CONCURRENT=1,2,4,8 python profiles/c5_fit_sweep.py ...   # restarts in flight per GPU (default: the library's)
    torchrun --nproc-per-node 8 profiles/c5_fit_sweep.py [n ...]  # 4 restarts per GPU
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch.distributed as dist

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import bayex_b200 as bayex
from bayex_b200 import fit_sweep

sizes = [int(a) for a in sys.argv[1:]] or [1024, 4096, 8192]
d, R, steps = 8, int(os.environ.get("BX_C5_RESTARTS", "32")), 100
conc = [int(c) for c in os.environ["BX_C5_CONCURRENT"].split(",")] if "BX_C5_CONCURRENT" in os.environ else [None]
out = {}
for n, C in [(n, c) for n in sizes for c in conc]:
    rng = np.random.default_rng(5)
    X = rng.uniform(0, 1, (n, d)).astype(np.float32)
    centres = rng.uniform(0.2, 0.8, (3, d))
    Y = sum(np.exp(-np.sum((X - c) ** 2, axis=1) / 0.1) for c in centres) + 0.01 * rng.standard_normal(n)
    raws = np.concatenate([rng.uniform(-8, -2, (R, 1)), rng.uniform(-1, 1, (R, 1 + d))], axis=1).astype(np.float32)
    fit_sweep.fit_restarts(X, Y.astype(np.float32), raws[:world * (C or 4)], trainsteps=2, concurrent=C)  # warm-up: allocations, graph capture
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    best, idx, table = fit_sweep.fit_restarts(X, Y.astype(np.float32), raws, trainsteps=steps, concurrent=C)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dt = time.perf_counter() - t0
    if rank == 0:
        flops = R * steps * (float(n) ** 3 + float(n) * n * (3 * d + 4))  # F_step ~ n^3
        out[f"n{n}_c{C}"] = dict(n_gpus=world, concurrent=C or fit_sweep.concurrent_restarts(n), restarts=R, steps=steps, sweep_s=dt, restarts_per_s=R / dt, fit_steps_per_s=R * steps / dt,
                            algorithmic_tflops=flops / dt / 1e12, best_restart=idx, best_loss=float(table[idx, 0]),
                            finite_losses=int(np.isfinite(table[:, 0]).sum()))
        print(n, out[f"n{n}_c{C}"], flush=True)
if rank == 0:
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open(f"gpurun_out/c5_fit_sweep_{world}gpu.json", "w"), indent=1)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
