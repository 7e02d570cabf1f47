"""world_size-2 gloo run of the sharded-sample host logic (candidate-axis sharding, packed max all-reduce, top-k
all-gather + merge).  Scores come from the oracle here (CPU box); on the GPU box the same helpers carry kernel output."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import ctypes
    from bayex_b200 import _lib, optimizer as bopt
    from oracle import acq as oacq, domain as odom, gp as ogp, prng
    rng = np.random.default_rng(3)
    X = rng.uniform(0, 1, (40, 3)).astype(np.float32)
    Y = np.sin(X.sum(1)).astype(np.float32)
    P = ogp.GPParams(-5.0, 0.2, np.array([0.1, 0.3, -0.2], np.float32))
    space = odom.ParamSpace({"a": odom.Real(0, 1), "b": odom.Real(0, 1), "c": odom.Real(0, 1)})
    f = ogp.Factor(P, X, Y, np.ones(40), np.float32)
    size, k = 3001, 50
    b, c = bopt.shard_range(size, rank, world)
    cand = space.to_array(space.sample_params(prng.key(9), (c,), offset=b))
    vals = oacq.expected_improvement(cand, X, Y, np.ones(40), P, factor=f)
    j = int(np.argmax(vals))
    packed = _lib.lib.bx_pack(ctypes.c_float(float(vals[j])), b + j)
    best = torch.tensor([packed - (1 << 64) if packed >= (1 << 63) else packed], dtype=torch.int64)
    best = bopt.allreduce_packed_max(best, dist)
    loc = np.argsort(vals + 0.0, kind="stable")[-k:]
    gv, gi = bopt.allgather_topk(torch.from_numpy(vals[loc]), torch.from_numpy((loc + b).astype(np.int32)), k, dist, world)
    tv, ti = bopt.merge_topk(gv, gi, k)
    q.put((rank, int(best.item()) & 0xFFFFFFFFFFFFFFFF, ti.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_argmax_and_topk_match_single_process():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    sys.path.insert(0, ROOT)
    from bayex_b200 import _lib
    from oracle import acq as oacq, domain as odom, gp as ogp, prng
    rng = np.random.default_rng(3)
    X = rng.uniform(0, 1, (40, 3)).astype(np.float32)
    Y = np.sin(X.sum(1)).astype(np.float32)
    P = ogp.GPParams(-5.0, 0.2, np.array([0.1, 0.3, -0.2], np.float32))
    space = odom.ParamSpace({"a": odom.Real(0, 1), "b": odom.Real(0, 1), "c": odom.Real(0, 1)})
    cand = space.to_array(space.sample_params(prng.key(9), (3001,)))
    vals = oacq.expected_improvement(cand, X, Y, np.ones(40), P)
    want_idx = int(np.argmax(vals))
    want_top = np.argsort(vals + 0.0, kind="stable")[-50:].tolist()
    for rank, packed, top in got:
        v, i = _lib.unpack(packed)
        assert i == want_idx and np.float32(v) == vals[want_idx]
        assert top == want_top


# ---- restart sweep (BASELINE config 5): round-robin sharding + all-gather of (loss, raw hypers) rows + arg-min ------
def _sweep_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from bayex_b200 import fit_sweep
    R, width = 7, 5
    losses = np.array([3.0, np.nan, 1.5, 1.5, np.inf, 2.0, 9.0], np.float32)  # tie at 2/3 -> first wins; NaN/inf never
    rows = {i: np.concatenate([[losses[i]], np.full(width - 1, 10.0 * i, np.float32)]) for i in fit_sweep.shard_restarts(R, rank, world)}
    table = fit_sweep.gather_restart_rows(rows, R, width, dist)
    q.put((rank, table.tolist(), fit_sweep.argmin_loss(table[:, 0])))
    dist.barrier()
    dist.destroy_process_group()


def test_restart_sweep_gather_and_argmin_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_sweep_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    sys.path.insert(0, ROOT)
    from bayex_b200 import fit_sweep
    assert fit_sweep.shard_restarts(7, 0, 2) == [0, 2, 4, 6] and fit_sweep.shard_restarts(7, 1, 2) == [1, 3, 5]
    assert sorted(fit_sweep.shard_restarts(32, r, 8)[0] for r in range(8)) == list(range(8))
    assert all(len(fit_sweep.shard_restarts(32, r, 8)) == 4 for r in range(8))
    tables = [np.array(t, np.float32) for _, t, _ in got]
    assert np.array_equal(tables[0], tables[1], equal_nan=True)          # every rank ends with the same table
    assert np.array_equal(tables[0][:, 1], 10.0 * np.arange(7))           # each row came from its owner
    assert [b for _, _, b in got] == [2, 2]
    single = fit_sweep.gather_restart_rows({i: tables[0][i] for i in range(7)}, 7, 5, None)
    assert np.array_equal(single, tables[0], equal_nan=True)
    import pytest
    with pytest.raises(FloatingPointError):
        fit_sweep.argmin_loss(np.array([np.nan, np.inf], np.float32))
