#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "nlml or posterior_fit or config5 or golden or readme or optim" > gpurun_out/pytest_n.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -5 gpurun_out/pytest_n.log
for n in 512 4096; do python profiles/fit_step_trace.py $n > gpurun_out/fit_trace_$n.csv 2> gpurun_out/fit_trace_$n.err; echo "trace $n rc=$?"; head -8 gpurun_out/fit_trace_$n.csv; done
python profiles/profile_paths.py > gpurun_out/profile_paths_n.log 2>&1; grep -E "^(512|1024|2048|4096|8192)|^c" gpurun_out/profile_paths_n.log | cut -c1-150
