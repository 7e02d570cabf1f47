#!/bin/bash
# round 2, 8 GPUs: bench line at N = 8 (no CPU leg: the other ranks would idle), multi-GPU parity test
set -u
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --no-cpu-baseline > gpurun_out/r2e_bench_n8.json 2> gpurun_out/r2e_bench_n8.err; echo "bench n8 rc=$?"; cut -c1-400 gpurun_out/r2e_bench_n8.json; tail -2 gpurun_out/r2e_bench_n8.err
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q > gpurun_out/r2e_pytest_mgpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2e_pytest_mgpu.log
