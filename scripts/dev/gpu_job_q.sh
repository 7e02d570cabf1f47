#!/bin/bash
# GPU job Q: shared-generation engine - parity, ablations, bench.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tcgen05_engine" > gpurun_out/pytest_q.log 2>&1; rc=$?; echo "engine test rc=$rc"; tail -12 gpurun_out/pytest_q.log
if [ $rc -ne 0 ]; then exit 1; fi
export ENGINE=tcgen05_f16_shared
source profiles/ablate.sh > gpurun_out/ablate_q.log 2>&1; cat gpurun_out/ablate_q.log
for e in tcgen05_f16_shared tcgen05_f16_pair; do
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --engine $e > gpurun_out/bench_q_$e.json 2> gpurun_out/bench_q_$e.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_q_$e.json").read().strip().splitlines()[-1])
print("$e value", round(d["value"]), "ms/step", round(d["ms_per_step"],1), "kernel_ms", round(d["roofline"]["kernel_ms"],1), "frac", round(d["roofline"]["frac"],4), d["clocks"])
PY
done
