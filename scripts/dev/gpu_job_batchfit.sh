#!/bin/bash
# batched restarts (bx_fit_adamw_batch): parity tests, then the sweep time against the number of restarts in flight
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "restart or side_by_side or posterior_fit" 2>&1 | tail -5
BX_C5_CONCURRENT=1,2,4,8 timeout 200 python profiles/c5_fit_sweep.py 1024 2>&1 | grep -E "^[0-9]+ " | cut -c1-200
BX_C5_RESTARTS=8 BX_C5_CONCURRENT=1,2,4,8 timeout 200 python profiles/c5_fit_sweep.py 4096 2>&1 | grep -E "^[0-9]+ " | cut -c1-200
BX_C5_RESTARTS=4 BX_C5_CONCURRENT=1,4 timeout 200 python profiles/c5_fit_sweep.py 8192 2>&1 | grep -E "^[0-9]+ " | cut -c1-200
