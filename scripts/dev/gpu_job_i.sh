#!/bin/bash
# GPU job I (2 GPUs): sharded bench through torchrun + the new parity test on the production engine.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "production_engine or tcgen05_engine" > gpurun_out/pytest_i.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_i.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_i_2gpu.json 2> gpurun_out/bench_i_2gpu.err; echo "bench2 rc=$?"; cut -c1-330 gpurun_out/bench_i_2gpu.json | tail -2
