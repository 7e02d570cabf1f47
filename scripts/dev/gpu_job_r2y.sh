#!/bin/bash
# r2y: incremental refit (bx_factor_append) and the streamed trmv kernels - parity tests of everything on the fit path,
# appended against rebuilt factors, timings
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -k "append or kept_hypers or score_points_vs_oracle or nlml_and_gradient or posterior_fit_vs_oracle or golden or failed_cholesky or state_rebuilds or adamw_kernel or restarts_side or config5 or readme or reference_test" > gpurun_out/r2y_tests.txt 2>&1
echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2y_tests.txt | tail -12
timeout 600 python scripts/dev/diag_append.py > gpurun_out/r2y_diag_append.txt 2>&1
echo "diag rc=$?"; tail -12 gpurun_out/r2y_diag_append.txt
timeout 600 python scripts/dev/append_time.py > gpurun_out/r2y_append_time.txt 2>&1
echo "append_time rc=$?"; cat gpurun_out/r2y_append_time.txt | tail -20
timeout 600 python profiles/profile_paths.py > gpurun_out/r2y_paths.txt 2>&1
echo "paths rc=$?"; cut -c1-220 gpurun_out/r2y_paths.txt | tail -12
