#!/bin/bash
# r2z2: the final build on 2 GPUs - multi-GPU parity test (incl. fit(trainsteps=0) on every rank) and a short 2-GPU bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multi_gpu.py -q > gpurun_out/r2z2_mgpu_tests.txt 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2z2_mgpu_tests.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 3 --warmup 3 --skip-c5 > gpurun_out/r2z2_bench_2gpu.json 2> gpurun_out/r2z2_bench_2gpu.err
echo "bench rc=$?"; cut -c1-300 gpurun_out/r2z2_bench_2gpu.json; tail -3 gpurun_out/r2z2_bench_2gpu.err
