#!/bin/bash
# final state of the round: smoke, the whole GPU suite, a bench line (restart sweep skipped: unchanged since r2z), small paths
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2zz_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2zz_smoke.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2zz_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2zz_pytest.log
python bench.py --skip-c5 > gpurun_out/r2zz_bench_1gpu.json 2> gpurun_out/r2zz_bench.err; echo "bench rc=$?"; cut -c1-260 gpurun_out/r2zz_bench_1gpu.json
timeout 300 python profiles/profile_paths.py > gpurun_out/r2zz_paths.txt 2>&1; echo "paths rc=$?"; cut -c1-200 gpurun_out/r2zz_paths.txt | tail -10
