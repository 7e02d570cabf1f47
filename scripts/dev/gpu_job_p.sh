#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "single_observation or more_than_32 or failed_cholesky or state_file" > gpurun_out/pytest_p.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -40 gpurun_out/pytest_p.log
