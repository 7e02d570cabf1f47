#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "nlml or posterior_fit or config5 or golden" > gpurun_out/pytest_y.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_y.log
for n in 512 2048 4096; do python profiles/fit_step_trace.py $n 2>/dev/null | head -3 | tr '\n' ' '; echo; done
CMD="python profiles/fit_step.py 8192"
$CMD > gpurun_out/plain_y.log 2>&1 && \
ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "bx_fit/" -k regex:gemm128_kernel -c 40 -o gpurun_out/prof_gemm128_r1d -f $CMD > gpurun_out/ncu_y.log 2>&1
echo "ncu rc=$?"
