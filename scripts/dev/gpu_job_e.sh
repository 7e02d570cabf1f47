#!/bin/bash
# GPU job E: evidence for the pair engine - bench line, launch list, full ncu capture at the bench size, non-headline paths.
set -u
mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_e.json 2> gpurun_out/bench_e.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/bench_e.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_e_ref.json 2> gpurun_out/bench_e_ref.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/bench_e_ref.json
python profiles/profile_paths.py > gpurun_out/profile_paths_e.log 2>&1; echo "paths rc=$?"; tail -12 gpurun_out/profile_paths_e.log
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_e1.log 2>&1 && \
ncu --nvtx --nvtx-include "bx_timed/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1c.csv $CMD > gpurun_out/ncu_e1.log 2>&1
echo "ncu launches rc=$?"
CMD2="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
$CMD2 > gpurun_out/plain_e2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:score_tc16p -s 1 -c 1 -o gpurun_out/prof_score_tc16p_full_r1 -f $CMD2 > gpurun_out/ncu_e2.log 2>&1
echo "ncu full rc=$?"
