#!/bin/bash
# round 2: blocked refinement products + 4-row-block SIMT engine: parity suite, A/B of the front half / refinement / C2 / C4
set -u
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2d_pytest.log | tail -20
timeout 900 python scripts/dev/ab_front.py > gpurun_out/r2d_ab_front.log 2>&1; echo "ab rc=$?"; cat gpurun_out/r2d_ab_front.log
