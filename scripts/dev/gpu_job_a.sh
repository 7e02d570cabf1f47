#!/bin/bash
# GPU job A: full parity suite, bench with AUTO (= fp16 engine), ablations of the fp16 engine, ncu launch list + full capture.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_a.log 2>&1; echo "pytest rc=$?" ; tail -3 gpurun_out/pytest_a.log
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/bench_a.json
export ENGINE=tcgen05_f16
source profiles/ablate.sh > gpurun_out/ablate_f16.log 2>&1; cat gpurun_out/ablate_f16.log
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_a.log 2>&1 && \
ncu --nvtx --nvtx-include "bx_timed/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1b.csv $CMD > gpurun_out/ncu_a.log 2>&1
echo "ncu launches rc=$?"
CMD2="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --candidates 2097152"
$CMD2 > gpurun_out/plain_b.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:score_tc16 -s 1 -c 1 -o gpurun_out/prof_score_tc16_r1 -f $CMD2 > gpurun_out/ncu_b.log 2>&1
echo "ncu full rc=$?"
