#!/bin/bash
# round 2, first validation of the two-tier path / fused top-k / device-side sample(): smoke, GPU suite, TMEM bandwidth
set -u
mkdir -p gpurun_out

timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2a_smoke.log
timeout 2400 python -m pytest tests -m gpu -q --deselect tests/test_gpu_parity.py::test_config3_full_size_properties > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2a_pytest.log | tail -40
