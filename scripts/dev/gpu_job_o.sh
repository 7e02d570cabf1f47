#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "nlml or posterior_fit or config5 or golden or readme or optim" > gpurun_out/pytest_o.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -3 gpurun_out/pytest_o.log
for pw in 1 2 4; do
  export BX_CHOL_PW=$pw
  BX_CHOL_PW=$pw timeout 300 python -m pytest tests -m gpu -x -q -k "nlml_and_gradient" 2>&1 | tail -1
  for n in 1024 4096 8192; do python profiles/fit_step_trace.py $n > gpurun_out/fit_trace_pw${pw}_$n.csv 2>/dev/null; echo "PW=$pw n=$n: $(head -1 gpurun_out/fit_trace_pw${pw}_$n.csv)"; done
done
head -9 gpurun_out/fit_trace_pw2_4096.csv
