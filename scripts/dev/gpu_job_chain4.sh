#!/bin/bash
# potrf_solve_kernel(next = 1): the first column of a panel also updates the second one
set -u
timeout 60 bayex_b200/csrc/build/bench_potrf_solve
timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "config5 or posterior_fit or nlml or golden or mixed_domain or state_rebuilds or failed_cholesky or side_by_side or readme or 1d_optim" 2>&1 | tail -3
timeout 200 python profiles/fit_step_time.py 128 512 1024 2048 4096 8192 2>&1 | grep "^n="
BX_CHOL_FUSE_NEXT=0 timeout 200 python profiles/fit_step_time.py 512 2048 2>&1 | grep "^n=" | sed 's/$/ fuse_next=0/'
