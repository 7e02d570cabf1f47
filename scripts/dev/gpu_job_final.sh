#!/bin/bash
# Final evidence of the round: bench line, reference arm, launch list, full ncu capture at the bench size, path timings.
set -u
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_f.json 2> gpurun_out/bench_f.err; echo "bench rc=$?"; cut -c1-220 gpurun_out/bench_f.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_f_ref.json 2> gpurun_out/bench_f_ref.err; echo "ref rc=$?"
python profiles/profile_paths.py > gpurun_out/profile_paths_f.log 2>&1; grep -E "^(23|64|512|1024|2048|4096|8192) |^c" gpurun_out/profile_paths_f.log | cut -c1-170
for n in 512 1024 2048 4096 8192; do python profiles/fit_step_trace.py $n > gpurun_out/fit_trace_$n.csv 2>/dev/null; head -1 gpurun_out/fit_trace_$n.csv; done
python profiles/sample_trace.py > gpurun_out/sample_trace.csv 2>/dev/null; grep "^#" gpurun_out/sample_trace.csv
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_f1.log 2>&1 && \
ncu --nvtx --nvtx-include "bx_timed/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1d.csv $CMD > gpurun_out/ncu_f1.log 2>&1
echo "ncu launches rc=$?"
CMD2="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
$CMD2 > gpurun_out/plain_f2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:score_tc16p -s 1 -c 1 -o gpurun_out/prof_score_tc16p_r1d -f $CMD2 > gpurun_out/ncu_f2.log 2>&1
echo "ncu full rc=$?"
