#!/bin/bash
# r2x: expanded-form test, ablations of the new kernel (K* math off / MMA off), bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "expanded_distance or production_path or config3 or config4" > gpurun_out/r2x_tests.txt 2>&1
tail -4 gpurun_out/r2x_tests.txt
for ab in 0 1 2; do
  echo "== BX_TC_ABLATE=$ab (1: K* math skipped, 2: MMA issue skipped)" >> gpurun_out/r2x_ablate.txt
  BX_TC_ABLATE=$ab timeout 300 python scripts/dev/front_only_time.py >> gpurun_out/r2x_ablate.txt 2>&1
done
cat gpurun_out/r2x_ablate.txt
timeout 900 python bench.py > gpurun_out/r2x_bench.json 2> gpurun_out/r2x_bench.err
tail -c 3000 gpurun_out/r2x_bench.json
