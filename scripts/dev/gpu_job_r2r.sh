#!/bin/bash
# round 2: accurate mode with the fold specialised into straight-line cases: accuracy / speed / ablation, suite, path timings
set -u
mkdir -p gpurun_out
timeout 600 python scripts/dev/check_acc.py 2>&1 | grep -E "^n=|ACC|RAW" > gpurun_out/r2r_check_acc.log; cat gpurun_out/r2r_check_acc.log
timeout 600 python scripts/dev/acc_ablate.py > gpurun_out/r2r_acc_ablate.log 2>&1; tail -4 gpurun_out/r2r_acc_ablate.log
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r2r_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2r_pytest.log | tail -12
timeout 900 python profiles/profile_paths.py 2>&1 | grep -E "^c" | cut -c1-200 > gpurun_out/r2r_paths.log; cat gpurun_out/r2r_paths.log
