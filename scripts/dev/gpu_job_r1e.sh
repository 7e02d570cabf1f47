#!/bin/bash
# r1e evidence: whole GPU suite + bench on the build with the fused Cholesky chain / batched restarts, fit-path traces, C5 sweep on one GPU
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r1e.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_r1e.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_r1e.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_r1e.log
python bench.py > gpurun_out/bench_r1e.json 2> gpurun_out/bench_r1e.err; echo "bench rc=$?"; cut -c1-250 gpurun_out/bench_r1e.json
python profiles/profile_paths.py > gpurun_out/profile_paths_r1e.log 2>&1; grep -E "^(23|64|512|1024|2048|4096|8192) |^c" gpurun_out/profile_paths_r1e.log | cut -c1-170
for n in 512 2048 4096; do python profiles/fit_step_trace.py $n > gpurun_out/fit_trace_r1e_$n.csv 2>/dev/null; head -1 gpurun_out/fit_trace_r1e_$n.csv; done
timeout 300 python profiles/c5_fit_sweep.py 1024 4096 8192 2>&1 | grep -E "^[0-9]+ " | cut -c1-200
