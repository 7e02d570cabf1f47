#!/bin/bash
# BASELINE config 5 on 8 GPUs: 32 restarts x 100 steps, 4 restarts per GPU fitted side by side
set -u
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 profiles/c5_fit_sweep.py 1024 4096 8192 2>&1 | grep -E "^[0-9]+ " | cut -c1-260
