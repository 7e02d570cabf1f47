#!/bin/bash
# round 2, 2 GPUs: full GPU suite incl. the multi-GPU parity test after the single-pass decision / weighted shards, C2 / C4 timings
set -u
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r2p_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2p_pytest.log | tail -12
timeout 900 python profiles/profile_paths.py 2>&1 | grep -E "^c" | cut -c1-200 > gpurun_out/r2p_paths.log; cat gpurun_out/r2p_paths.log
