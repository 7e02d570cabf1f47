#!/bin/bash
# r1f: final validation of the round's last build: smoke, whole GPU suite, bench (both arms), path timings, fit traces
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r1f.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_r1f.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_r1f.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_r1f.log
python bench.py > gpurun_out/bench_r1f.json 2> gpurun_out/bench_r1f.err; echo "bench rc=$?"; cut -c1-250 gpurun_out/bench_r1f.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r1f_ref.json 2> gpurun_out/bench_r1f_ref.err; echo "ref rc=$?"; cut -c1-120 gpurun_out/bench_r1f_ref.json
python profiles/profile_paths.py > gpurun_out/profile_paths_r1f.log 2>&1; grep -E "^(23|64|512|1024|2048|4096|8192) |^c" gpurun_out/profile_paths_r1f.log | cut -c1-170
for n in 512 2048 4096; do python profiles/fit_step_trace.py $n > gpurun_out/fit_trace_r1f_$n.csv 2>/dev/null; head -1 gpurun_out/fit_trace_r1f_$n.csv; done
timeout 100 python scripts/dev/fit_call_time.py 2>&1 | grep n0=
