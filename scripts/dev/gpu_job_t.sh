#!/bin/bash
set -u
mkdir -p gpurun_out
for n in 512 2048; do python profiles/fit_step_trace.py $n > gpurun_out/fit_trace_$n.csv 2>/dev/null; head -14 gpurun_out/fit_trace_$n.csv; done
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --candidates 2097152 2>/dev/null | cut -c1-200
