#!/bin/bash
# Full validation on one B200: smoke, the whole GPU parity suite, the bench line (ours + reference arm), non-headline paths.
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_final.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_final.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_final.log
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"; cut -c1-250 gpurun_out/bench_final.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_final_ref.json 2> gpurun_out/bench_final_ref.err; echo "ref rc=$?"; cut -c1-200 gpurun_out/bench_final_ref.json
python profiles/profile_paths.py > gpurun_out/profile_paths_final.log 2>&1; grep -E "^(23|64|512|1024|2048|4096|8192) |^c" gpurun_out/profile_paths_final.log | cut -c1-170
