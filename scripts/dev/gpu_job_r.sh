#!/bin/bash
# GPU job R: ring-depth variants of the pair engine at the bench size (kernel time only).
set -u
mkdir -p gpurun_out
V=bayex_b200/csrc/build/variants
for name in base a4b2u3 a4b2u2 a3b2u3 a3b3u2; do
  if [ "$name" == "base" ]; then unset BAYEX_B200_LIB; else export BAYEX_B200_LIB=$PWD/$V/$name.so; fi
  timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_r_$name.json 2> gpurun_out/bench_r_$name.err; rc=$?
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_r_$name.json").read().strip().splitlines()[-1])
    print("$name rc=$rc kernel_ms", round(d["roofline"]["kernel_ms"],1), "ms/step", round(d["ms_per_step"],1), d["clocks"]["sm_mhz"])
except Exception as e:
    print("$name rc=$rc failed", e)
PY
done
