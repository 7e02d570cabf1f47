#!/bin/bash
# fused potrf + panel solve (potrf_solve_kernel): fit-path parity, then per-step time against the unfused chain
set -u
timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "config5 or posterior_fit or nlml or golden or mixed_domain or state_rebuilds or failed_cholesky or side_by_side" 2>&1 | tail -5
BX_CHOL_FUSED=0 timeout 200 python profiles/fit_step_time.py 128 512 1024 2048 4096 8192 2>&1 | grep "^n=" | sed 's/$/ unfused/'
timeout 200 python profiles/fit_step_time.py 128 512 1024 2048 4096 8192 2>&1 | grep "^n="
