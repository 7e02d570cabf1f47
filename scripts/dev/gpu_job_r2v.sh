#!/bin/bash
# r2v: refinement loop: programmatic dependent launch + 16-byte staging of the T blocks: timing + tests
mkdir -p gpurun_out
echo "== BX_NO_PDL=1" > gpurun_out/r2v_refine.txt
BX_NO_PDL=1 timeout 120 python scripts/dev/refine_time.py 7 16 50 >> gpurun_out/r2v_refine.txt 2>&1
echo "== PDL" >> gpurun_out/r2v_refine.txt
timeout 120 python scripts/dev/refine_time.py 7 16 50 >> gpurun_out/r2v_refine.txt 2>&1
echo "rc=$?" >> gpurun_out/r2v_refine.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "refine or sample or config" > gpurun_out/r2v_tests.txt 2>&1
cat gpurun_out/r2v_refine.txt; tail -5 gpurun_out/r2v_tests.txt
