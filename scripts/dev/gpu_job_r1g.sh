#!/bin/bash
# r1g: bench line + smoke of the round's final build (the GPU suite of this build: gpurun_out/pytest_r1g.log, 72 passed)
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r1g.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_r1g.log
python bench.py > gpurun_out/bench_r1g.json 2> gpurun_out/bench_r1g.err; echo "bench rc=$?"; cut -c1-250 gpurun_out/bench_r1g.json
python profiles/profile_paths.py > gpurun_out/profile_paths_r1g.log 2>&1; grep -E "^(23|64|512|1024|2048|4096|8192) |^c" gpurun_out/profile_paths_r1g.log | cut -c1-170
