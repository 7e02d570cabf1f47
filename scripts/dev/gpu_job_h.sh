#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "nlml or posterior_fit or config5 or golden or readme or optim" > gpurun_out/pytest_h.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -5 gpurun_out/pytest_h.log
python profiles/profile_paths.py > gpurun_out/profile_paths_h.log 2>&1; echo "paths rc=$?"; grep -E "^(512|1024|2048|4096|8192)" gpurun_out/profile_paths_h.log
CMD="python profiles/fit_step.py 4096"
$CMD > gpurun_out/plain_h.log 2>&1 && \
ncu --nvtx --nvtx-include "bx_fit/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_fit_4096.csv $CMD > gpurun_out/ncu_h.log 2>&1
echo "ncu fit rc=$?"
