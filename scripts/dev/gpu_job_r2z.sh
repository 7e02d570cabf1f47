#!/bin/bash
# r2z: final evidence of the round - smoke, the whole GPU suite, bench lines (ours + reference arm), launch list of the timed
# bench region and `ncu --set full` of the scoring kernel at the bench size (the build with the expanded distance form)
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2z_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2z_smoke.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2z_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2z_pytest.log
python bench.py > gpurun_out/r2z_bench_1gpu.json 2> gpurun_out/r2z_bench.err; echo "bench rc=$?"; cut -c1-260 gpurun_out/r2z_bench_1gpu.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2z_bench_reference_arm.json 2> gpurun_out/r2z_bench_ref.err; echo "ref rc=$?"; cut -c1-200 gpurun_out/r2z_bench_reference_arm.json
B="python bench.py --steps 2 --warmup 3 --skip-fit --no-cpu-baseline"
$B > gpurun_out/r2z_bench_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "bx_timed/" --csv --log-file gpurun_out/r2z_launches_bench.csv $B > gpurun_out/r2z_bench_ncu.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/r2z_launches_bench.csv
F="python scripts/dev/front_only.py"
$F > gpurun_out/r2z_front_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:score_tc16p_kernel -s 2 -c 1 -o gpurun_out/r2z_score_tc16p $F > gpurun_out/r2z_front_ncu.log 2>&1
echo "score ncu rc=$?"; tail -2 gpurun_out/r2z_front_ncu.log
ls -la gpurun_out/*.ncu-rep
