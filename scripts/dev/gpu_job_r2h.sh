#!/bin/bash
# round 2: chunk length of the accurate mode (1 / 2 / 4 K-blocks per fold): accuracy, speed, ablations; refinement after the
# staging fix
set -u
mkdir -p gpurun_out
V=bayex_b200/csrc/build/variants
for lib in default $V/acc_chj2.so $V/acc_chj4.so; do
  if [ "$lib" != "default" ]; then export BAYEX_B200_LIB=$PWD/$lib; fi
  echo "=== $lib"
  timeout 600 python scripts/dev/check_acc.py 2>&1 | grep -E "^n=|ACC|SIMT"
  timeout 600 python scripts/dev/acc_ablate.py 2>&1 | tail -4
done > gpurun_out/r2h_acc_chunks.log 2>&1
unset BAYEX_B200_LIB
cat gpurun_out/r2h_acc_chunks.log
timeout 600 python profiles/sample_trace_c3.py > gpurun_out/r2h_sample_trace_c3.log 2>&1; echo "trace rc=$?"; grep -E "refine|sample\(\)|sample_front|bx::" gpurun_out/r2h_sample_trace_c3.log
