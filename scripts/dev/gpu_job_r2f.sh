#!/bin/bash
# round 2: accurate tensor mode (accuracy + speed against SIMT / fast mode), sample() trace at one rank's share of C3, then the
# parity suite with the accurate mode as tier 2
set -u
mkdir -p gpurun_out
timeout 900 python scripts/dev/check_acc.py > gpurun_out/r2f_check_acc.log 2>&1; echo "check_acc rc=$?"; cat gpurun_out/r2f_check_acc.log
timeout 600 python profiles/sample_trace_c3.py > gpurun_out/r2f_sample_trace_c3.log 2>&1; echo "trace rc=$?"; cat gpurun_out/r2f_sample_trace_c3.log
BX_TIER2=acc timeout 2400 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_parity.py::test_config3_full_size_properties > gpurun_out/r2f_pytest_acc.log 2>&1; echo "pytest(acc) rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed|Error" gpurun_out/r2f_pytest_acc.log | tail -20
