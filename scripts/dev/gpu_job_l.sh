#!/bin/bash
# GPU job L (8 GPUs): sharded bench + config-5 restart sweep.
set -u
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo "bench8 rc=$?"; tail -1 gpurun_out/bench_n8.json | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 profiles/c5_fit_sweep.py 1024 4096 8192 > gpurun_out/c5_8gpu_final.log 2>&1; echo "c5 rc=$?"; grep -E "^(1024|4096|8192) " gpurun_out/c5_8gpu_final.log
