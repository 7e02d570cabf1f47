#!/bin/bash
# round 2, 8 GPUs: per-phase device timeline of sample() on rank 0 (BX_TRACE_SAMPLE), bench line with weighted shards
set -u
mkdir -p gpurun_out
BX_TRACE_SAMPLE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --no-cpu-baseline --skip-fit --steps 4 > gpurun_out/r2o_bench_n8_trace.log 2> gpurun_out/r2o_bench_n8_trace.err; echo "rc=$?"
grep "device timeline" gpurun_out/r2o_bench_n8_trace.log | tail -6
grep -o '"value": [0-9.]*' gpurun_out/r2o_bench_n8_trace.log | head -2
timeout 600 python bench.py --no-cpu-baseline --skip-fit --steps 4 > gpurun_out/r2o_bench_n1.json 2> gpurun_out/r2o_bench_n1.err; grep -o '"value": [0-9.]*' gpurun_out/r2o_bench_n1.json | head -2
