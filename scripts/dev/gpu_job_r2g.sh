#!/bin/bash
# round 2: accurate tensor mode after the epilogue-sync fixes: accuracy / speed, A/B of the front half and C2 / C4, parity suite
set -u
mkdir -p gpurun_out
timeout 900 python scripts/dev/check_acc.py > gpurun_out/r2g_check_acc.log 2>&1; echo "check_acc rc=$?"; grep -E "^n=|ACC|RAW" gpurun_out/r2g_check_acc.log
timeout 900 python scripts/dev/ab_front.py > gpurun_out/r2g_ab_front.log 2>&1; echo "ab rc=$?"; cat gpurun_out/r2g_ab_front.log
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed|Error" gpurun_out/r2g_pytest.log | tail -20
