#!/bin/bash
# look-ahead Cholesky: parity of the fit path at n >= 2048, then per-step time with and without it
set -u
timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "config5 or posterior_fit or nlml or mixed_domain or state_rebuilds" 2>&1 | tail -5
BX_CHOL_LOOKAHEAD_MIN=0 timeout 200 python profiles/fit_step_time.py 2048 4096 8192 2>&1 | grep "^n="
timeout 200 python profiles/fit_step_time.py 2048 4096 8192 2>&1 | grep "^n="
BX_CHOL_LOOKAHEAD_MIN=512 timeout 200 python profiles/fit_step_time.py 512 1024 2>&1 | grep "^n="
timeout 200 python profiles/fit_step_time.py 512 1024 2>&1 | grep "^n="
