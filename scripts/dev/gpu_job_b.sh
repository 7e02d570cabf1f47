#!/bin/bash
# GPU job B: new fp16 pair engine - parity first, then bench + ablations.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tcgen05_engine" > gpurun_out/pytest_b1.log 2>&1; rc=$?; echo "engine test rc=$rc"; tail -15 gpurun_out/pytest_b1.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_b2.log 2>&1; echo "pytest rc=$?" ; tail -3 gpurun_out/pytest_b2.log
for e in tcgen05_f16_pair tcgen05_f16; do
  timeout 600 python bench.py --steps 3 --warmup 3 --engine $e --no-cpu-baseline > gpurun_out/bench_b_$e.json 2> gpurun_out/bench_b_$e.err; echo "bench $e rc=$?"
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_b_$e.json").read().strip().splitlines()[-1])
print("$e", "value", round(d["value"]), "ms/step", round(d["ms_per_step"],1), "kernel_ms", round(d["roofline"]["kernel_ms"],1), "frac", round(d["roofline"]["frac"],4), d["clocks"])
PY
done
export ENGINE=tcgen05_f16_pair
source profiles/ablate.sh > gpurun_out/ablate_f16_pair.log 2>&1; cat gpurun_out/ablate_f16_pair.log
