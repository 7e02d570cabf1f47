#!/bin/bash
# round 2: ncu --set full + source page of the accurate mode alone (2^19 candidates at C3's factor)
set -u
mkdir -p gpurun_out
F="python scripts/dev/front_only.py 524288 acc"
$F > gpurun_out/r2q_acc_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:score_tc16p_kernel -s 1 -c 1 -o gpurun_out/r2q_score_acc $F > gpurun_out/r2q_acc_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/r2q_acc_ncu.log
