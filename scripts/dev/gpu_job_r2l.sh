#!/bin/bash
# round 2 evidence: launch list of the timed bench region, ncu --set full of the scoring kernel (fast + accurate mode) at the
# bench size, ncu --set full of the fit path's GEMM / Cholesky-chain kernels at n = 4096
set -u
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --skip-fit --no-cpu-baseline"
$B > gpurun_out/r2l_bench_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "bx_timed/" --csv --log-file gpurun_out/r2l_launches_bench.csv $B > gpurun_out/r2l_bench_ncu.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/r2l_launches_bench.csv
F="python scripts/dev/front_only.py"
$F > gpurun_out/r2l_front_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:score_tc16p_kernel -s 2 -c 2 -o gpurun_out/r2l_score_tc16p $F > gpurun_out/r2l_front_ncu.log 2>&1
echo "score ncu rc=$?"; tail -2 gpurun_out/r2l_front_ncu.log
G="python profiles/fit_step.py 4096"
$G > gpurun_out/r2l_fit_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "bx_fit/" -k regex:"gemm128_kernel|potrf_solve_kernel" -c 8 -o gpurun_out/r2l_fit_kernels $G > gpurun_out/r2l_fit_ncu.log 2>&1
echo "fit ncu rc=$?"; tail -2 gpurun_out/r2l_fit_ncu.log
ls -la gpurun_out/*.ncu-rep
