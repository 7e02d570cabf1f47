#!/bin/bash
# round 2: full GPU suite (incl. the C3 full-size test), default bench line, reference arm
set -u
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -s > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed|tensor engine alone" gpurun_out/r2b_pytest.log | tail -40
timeout 900 python bench.py > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench rc=$?"; cut -c1-1500 gpurun_out/r2b_bench.json; tail -3 gpurun_out/r2b_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2b_bench_ref.json 2> gpurun_out/r2b_bench_ref.err; echo "ref rc=$?"; cut -c1-400 gpurun_out/r2b_bench_ref.json
