#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "refine or sample or readme or optim or config2 or config3 or production" > gpurun_out/pytest_j.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -5 gpurun_out/pytest_j.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --candidates 2097152 > gpurun_out/bench_j.json 2> gpurun_out/bench_j.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_j.json").read().strip().splitlines()[-1])
print("2^21: ms/step", round(d["ms_per_step"],2), "kernel_ms", round(d["roofline"]["kernel_ms"],2), "overhead", round(d["ms_per_step"]-d["roofline"]["kernel_ms"],2), "e2e ms", round(d["e2e"]["ms_per_step"],2))
PY
python profiles/profile_paths.py > gpurun_out/profile_paths_j.log 2>&1; grep -E "^c[124]" gpurun_out/profile_paths_j.log
