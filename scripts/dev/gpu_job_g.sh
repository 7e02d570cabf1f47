#!/bin/bash
set -u
mkdir -p gpurun_out
for n in 4096; do
CMD="python profiles/fit_step.py $n"
$CMD > gpurun_out/plain_g_$n.log 2>&1 && \
ncu --nvtx --nvtx-include "bx_fit/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_fit_$n.csv $CMD > gpurun_out/ncu_g_$n.log 2>&1
echo "ncu fit $n rc=$?"
done
