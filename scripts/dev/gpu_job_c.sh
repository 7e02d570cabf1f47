#!/bin/bash
# GPU job C: ncu full capture of the fp16 pair engine at 2^21 candidates.
set -u
mkdir -p gpurun_out
CMD2="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --candidates 2097152 --engine tcgen05_f16_pair"
$CMD2 > gpurun_out/plain_c.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:score_tc16p -s 1 -c 1 -o gpurun_out/prof_score_tc16p_r1 -f $CMD2 > gpurun_out/ncu_c.log 2>&1
echo "ncu full rc=$?"
