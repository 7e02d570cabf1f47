#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "config5" > gpurun_out/pytest_k.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -5 gpurun_out/pytest_k.log
timeout 600 python profiles/c5_fit_sweep.py 1024 4096 8192 > gpurun_out/c5_1gpu.log 2>&1; echo "c5 rc=$?"; tail -4 gpurun_out/c5_1gpu.log
