#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tcgen05_engine or production" > gpurun_out/pytest_s.log 2>&1; rc=$?; echo "engine test rc=$rc"; tail -3 gpurun_out/pytest_s.log
V=bayex_b200/csrc/build/variants
for name in base w0 w256; do
  if [ "$name" == "base" ]; then unset BAYEX_B200_LIB; else export BAYEX_B200_LIB=$PWD/$V/$name.so; fi
  timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_s_$name.json 2> gpurun_out/bench_s_$name.err; rc=$?
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_s_$name.json").read().strip().splitlines()[-1])
    print("$name rc=$rc kernel_ms", round(d["roofline"]["kernel_ms"],1), "ms/step", round(d["ms_per_step"],1), d["clocks"]["sm_mhz"])
except Exception as e:
    print("$name rc=$rc failed", e)
PY
done
