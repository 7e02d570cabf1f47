#!/bin/bash
# round 2: barrier waits with suspend-time hints instead of polling: A/B of the front half, accurate mode speed, suite, bench
set -u
mkdir -p gpurun_out
timeout 900 python scripts/dev/ab_front.py > gpurun_out/r2s_ab_front.log 2>&1; echo "ab rc=$?"; cat gpurun_out/r2s_ab_front.log
timeout 600 python scripts/dev/acc_ablate.py > gpurun_out/r2s_acc_ablate.log 2>&1; tail -4 gpurun_out/r2s_acc_ablate.log
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r2s_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2s_pytest.log | tail -12
timeout 900 python bench.py --skip-c5 > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/r2s_bench.json
