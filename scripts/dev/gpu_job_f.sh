#!/bin/bash
# GPU job F: fit-path GEMM/potrf changes - full parity suite, then the non-headline path timings.
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_f.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -15 gpurun_out/pytest_f.log
python profiles/profile_paths.py > gpurun_out/profile_paths_f.log 2>&1; echo "paths rc=$?"; tail -12 gpurun_out/profile_paths_f.log
