#!/bin/bash
# round 2, 2 GPUs: full GPU suite (skinny refinement, multi-GPU parity test at world 2), bench at N = 2, sample() trace
set -u
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2c_pytest.log | tail -20
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --skip-c5 > gpurun_out/r2c_bench_n2.json 2> gpurun_out/r2c_bench_n2.err; echo "bench n2 rc=$?"; cut -c1-700 gpurun_out/r2c_bench_n2.json; tail -3 gpurun_out/r2c_bench_n2.err
timeout 600 python bench.py --skip-fit --no-cpu-baseline > gpurun_out/r2c_bench_n1.json 2> gpurun_out/r2c_bench_n1.err; echo "bench n1 rc=$?"; cut -c1-300 gpurun_out/r2c_bench_n1.json
timeout 600 python scripts/dev/ab_front.py > gpurun_out/r2c_ab_front.log 2>&1; echo "ab rc=$?"; cat gpurun_out/r2c_ab_front.log
