#!/bin/bash
# GPU job D: quick iteration loop for the pair engine - parity of the engine test, ablations at 2^21, full bench.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tcgen05_engine or config3" > gpurun_out/pytest_d.log 2>&1; rc=$?; echo "engine test rc=$rc"; tail -4 gpurun_out/pytest_d.log
if [ $rc -ne 0 ]; then exit 1; fi
export ENGINE=tcgen05_f16_pair
source profiles/ablate.sh > gpurun_out/ablate_d.log 2>&1; cat gpurun_out/ablate_d.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_d.json 2> gpurun_out/bench_d.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_d.json").read().strip().splitlines()[-1])
print("value", round(d["value"]), "ms/step", round(d["ms_per_step"],1), "kernel_ms", round(d["roofline"]["kernel_ms"],1), "frac", round(d["roofline"]["frac"],4), d["clocks"])
PY
