#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "nlml or posterior_fit or config5 or golden or refine" > gpurun_out/pytest_u.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -3 gpurun_out/pytest_u.log
for n in 512 1024 2048 4096; do python profiles/fit_step_trace.py $n > gpurun_out/fit_trace_$n.csv 2>/dev/null; head -1 gpurun_out/fit_trace_$n.csv; done
head -9 gpurun_out/fit_trace_512.csv
BX_GEMM64=1 python profiles/fit_step_trace.py 512 | head -1
BX_GEMM64=1 python profiles/fit_step_trace.py 1024 | head -1
