#!/bin/bash
# round 2: compute-sanitizer memcheck over every kernel of the path (small sizes), timings of the non-headline paths
set -u
mkdir -p gpurun_out
timeout 600 python scripts/dev/sanity_small.py > gpurun_out/r2m_sanity_plain.log 2>&1; echo "plain rc=$?"; tail -2 gpurun_out/r2m_sanity_plain.log
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 99 python scripts/dev/sanity_small.py > gpurun_out/r2m_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|Invalid|sanity ok" gpurun_out/r2m_memcheck.log | head
timeout 900 python profiles/profile_paths.py > gpurun_out/r2m_profile_paths.log 2>&1; echo "paths rc=$?"; grep -E "^(23|64|512|1024|2048|4096|8192) |^c" gpurun_out/r2m_profile_paths.log | cut -c1-200
