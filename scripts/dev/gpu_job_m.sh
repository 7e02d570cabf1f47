#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "optimize_suggestion or refine" > gpurun_out/pytest_m.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_m.log
for n in 4096 8192; do python profiles/fit_step_trace.py $n > gpurun_out/fit_trace_$n.csv 2> gpurun_out/fit_trace_$n.err; echo "trace $n rc=$?"; head -12 gpurun_out/fit_trace_$n.csv; done
