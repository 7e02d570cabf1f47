#!/bin/bash
# two-column warp Cholesky + alpha/NLML beside K^-1: fit-path parity, per-step time, restart-sweep concurrency
set -u
timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "config5 or posterior_fit or nlml or golden or mixed_domain or state_rebuilds or failed_cholesky or side_by_side or readme or 1d_optim" 2>&1 | tail -5
timeout 200 python profiles/fit_step_time.py 128 512 1024 2048 4096 8192 2>&1 | grep "^n="
BX_NO_HELPER_STREAM=1 timeout 200 python profiles/fit_step_time.py 128 512 1024 2>&1 | grep "^n=" | sed 's/$/ no_helper_stream/'
BX_C5_CONCURRENT=8,16,32 timeout 200 python profiles/c5_fit_sweep.py 1024 2>&1 | grep -E "^[0-9]+ " | cut -c1-120
BX_C5_RESTARTS=16 BX_C5_CONCURRENT=8,16 timeout 200 python profiles/c5_fit_sweep.py 4096 2>&1 | grep -E "^[0-9]+ " | cut -c1-120
