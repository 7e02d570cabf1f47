#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "nlml or posterior_fit or config5 or golden or refine" > gpurun_out/pytest_v.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -2 gpurun_out/pytest_v.log
for wm in 296 0 100000; do
for n in 2048 4096 8192; do echo "WIDE_MAX=$wm $(BX_GEMM_WIDE_MAX=$wm python profiles/fit_step_trace.py $n 2>/dev/null | head -1)"; done; done
python profiles/fit_step_trace.py 4096 2>/dev/null | head -9
