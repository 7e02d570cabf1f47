#!/bin/bash
# r2zz-b: register-resident sweep fit kernel (n <= 128) - every test of the fit path at small n, per-step timing old / new
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "posterior_fit_vs_oracle or adamw_kernel or nlml_and_gradient or golden or readme or reference_test or restarts_side or single_observation or failed_cholesky or state_rebuilds or kept_hypers or append" > gpurun_out/r2zz_fit_tests.txt 2>&1
echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed|Error|assert" gpurun_out/r2zz_fit_tests.txt | tail -15
timeout 300 python scripts/dev/small_fit_time.py > gpurun_out/r2zz_small_fit_time.txt 2>&1; cat gpurun_out/r2zz_small_fit_time.txt | tail -10
echo "--- previous paths"
BX_NO_FIT_SWEEP=1 timeout 300 python scripts/dev/small_fit_time.py > gpurun_out/r2zz_small_fit_time_prev.txt 2>&1; cat gpurun_out/r2zz_small_fit_time_prev.txt | tail -10
