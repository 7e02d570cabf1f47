#!/bin/bash
# round 2, 8 GPUs: bench line at N = 8 with throughput-weighted shards (no CPU leg), multi-GPU parity test
set -u
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --no-cpu-baseline > gpurun_out/r2k_bench_n8.json 2> gpurun_out/r2k_bench_n8.err; echo "bench n8 rc=$?"; grep -o '"value": [0-9.]*' gpurun_out/r2k_bench_n8.json | head -2; grep -o '"kernel_ms_per_rank": [^]]*]' gpurun_out/r2k_bench_n8.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --no-cpu-baseline --skip-fit > gpurun_out/r2k_bench_n4.json 2> gpurun_out/r2k_bench_n4.err; echo "bench n4 rc=$?"; grep -o '"value": [0-9.]*' gpurun_out/r2k_bench_n4.json | head -2
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q > gpurun_out/r2k_pytest_mgpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2k_pytest_mgpu.log
