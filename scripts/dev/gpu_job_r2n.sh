#!/bin/bash
# round 2: refinement with the K* build folded into the update kernel: parity suite, refinement / small-path timings
set -u
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r2n_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2n_pytest.log | tail -12
timeout 900 python scripts/dev/ab_front.py 2>&1 | grep -E "refinement|sample\(\)" > gpurun_out/r2n_ab.log; cat gpurun_out/r2n_ab.log
timeout 900 python profiles/profile_paths.py 2>&1 | grep -E "^c" | cut -c1-200
