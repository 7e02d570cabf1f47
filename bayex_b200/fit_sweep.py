"""Hyper-parameter fit sweep over random restarts (BASELINE config 5).

Every restart is an independent ``posterior_fit`` (gp.py:90-110: 100 x {NLML gradient, clip 10, AdamW 1e-3}) from its own
raw hypers, so the restart axis shards with no data-path collective: rank r runs restarts r, r + world, ...; one
all-gather of ``(loss, raw hypers)`` rows follows and every rank takes the same arg-min (first occurrence, NaN losses
never win).  Gram build and Cholesky of one restart stay on one GPU (north star).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib, engine
from .gp import params_from_raw


def concurrent_restarts(n: int) -> int:
    """Restarts fitted side by side per GPU by default (workspace: ~2 np^2 floats each).  Measured on one B200 (32
    restarts x 100 steps): n = 1024 1.85 s one after the other -> 0.29 / 0.22 / 0.20 s at 8 / 16 / 32 in flight; n = 4096
    (8 restarts) 3.19 -> 1.58 s at 8, no further gain at 16; n = 8192 (4 restarts) 7.0 -> 5.3 s at 4."""
    return 32 if n <= 1024 else 16 if n <= 2048 else 8


def shard_restarts(n_restarts: int, rank: int, world: int):
    """Indices of the restarts owned by ``rank`` (round-robin: equal counts whenever world divides n_restarts)."""
    return list(range(rank, n_restarts, world))


def gather_restart_rows(local_rows: dict, n_restarts: int, width: int, dist=None, group=None, device="cpu") -> np.ndarray:
    """All-gather per-restart result rows ``{restart index: [loss, raw...]}`` into one (n_restarts, width) array.

    Every rank contributes a dense (n_restarts, width) buffer holding its own rows and NaN elsewhere; after the
    all-gather each row is taken from the rank that owns it (rank = index mod world), so the result is identical on
    every rank and independent of the backend's reduction order."""
    buf = torch.full((n_restarts, width), float("nan"), dtype=torch.float32, device=device)
    for i, row in local_rows.items():
        buf[i] = torch.as_tensor(np.asarray(row, dtype=np.float32), device=device)
    if dist is None:
        return buf.cpu().numpy()
    world = dist.get_world_size(group)
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf, group=group)
    out = np.empty((n_restarts, width), dtype=np.float32)
    host = [p.cpu().numpy() for p in parts]
    for i in range(n_restarts):
        out[i] = host[i % world][i]
    return out


def argmin_loss(losses: np.ndarray) -> int:
    """First occurrence of the smallest finite loss; a NaN / inf loss (failed Cholesky) never wins."""
    l = np.where(np.isfinite(losses), losses, np.inf)
    if not np.isfinite(l).any():
        raise FloatingPointError("every restart produced a non-finite loss")
    return int(np.argmin(l))


def fit_restarts(x, y, raws, lr: float = 1e-3, trainsteps: int = 100, concurrent: int = None):
    """Run ``posterior_fit`` from every row of ``raws`` ((R, d + 2): [noise, amplitude, lengthscale...] raw values) on
    the compacted observations (x (n, d), y (n,)), sharded over the ``torch.distributed`` ranks if a group is
    initialised.  A rank fits its restarts ``concurrent`` at a time side by side on its GPU (``bx_fit_adamw_batch``;
    default ``concurrent_restarts(n)``; 1 = one after the other) - the results do not depend on it.  Returns
    ``(best GPParams, best index, table)`` with ``table[r] = [loss_r, fitted raw_r...]``; the loss is
    ``neg_log_likelihood`` (gp.py:78-87) at the fitted hypers."""
    import torch.distributed as tdist
    raws = np.ascontiguousarray(raws, dtype=np.float32)
    R, P = raws.shape
    d = np.asarray(x).shape[1]
    assert P == d + 2, "raws must have d + 2 columns"
    concurrent = concurrent_restarts(np.asarray(x).shape[0]) if concurrent is None else int(concurrent)
    if not 1 <= concurrent <= _lib.FIT_BATCH_MAX:
        raise ValueError(f"concurrent must be in [1, {_lib.FIT_BATCH_MAX}]")
    dist = tdist if (tdist.is_available() and tdist.is_initialized() and tdist.get_world_size() > 1) else None
    rank = dist.get_rank() if dist else 0
    world = dist.get_world_size() if dist else 1
    rows = {}
    post = None
    mine = shard_restarts(R, rank, world)
    for c0 in range(0, len(mine), concurrent):
        chunk = mine[c0:c0 + concurrent]
        if post is None:
            post = engine.Posterior(x, y, raws[chunk[0]], want_tc=False)
        fitted, recs = post.fit_hypers_batch(raws[chunk], lr=lr, steps=trainsteps)
        for j, r in enumerate(chunk):
            rows[r] = np.concatenate([[recs[j, 1]], fitted[j]])
    dev = post.dev if post is not None else engine.device()
    table = gather_restart_rows(rows, R, P + 1, dist, device=dev if dist else "cpu")
    best = argmin_loss(table[:, 0])
    return params_from_raw(table[best, 1:], d), best, table
