"""Multi-GPU parity on real hardware: the candidate axis and the multi-start axis shard over the ranks, the
restart sweep shards by restart - and every rank must return exactly what one GPU returns.  Skips unless the box has at
least two GPUs (the driver's `pytest -m gpu` box has one; `gpurun --gpus 2` runs this)."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "mgpu_worker.py")


def _run(world, prefix):
    env = dict(os.environ, OMP_NUM_THREADS="1")
    if world == 1:
        cmd = [sys.executable, WORKER, prefix]
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
               "--master-port", str(29500 + world), WORKER, prefix]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    return [json.load(open(f"{prefix}.rank{k}.json")) for k in range(world)]


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_sample_and_sweep_equal_single_gpu(world, tmp_path):
    if not torch.cuda.is_available() or torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    one = _run(1, str(tmp_path / "one"))[0]
    many = _run(world, str(tmp_path / f"w{world}"))
    for r in many:
        for key, want in one.items():
            if key in ("world", "rank"):
                continue
            assert r[key] == want, f"rank {r['rank']} of {world}: {key} differs from the single-GPU run"
