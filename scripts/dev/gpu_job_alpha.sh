#!/bin/bash
# alpha refinement in the posterior build: whole GPU suite, then the failing case's margins with and without it
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_r1g.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_r1g.log
timeout 300 python scripts/dev/diag_mean.py 2>&1 | grep -E "^n="
BX_ALPHA_REFINE=0 timeout 300 python scripts/dev/diag_mean.py 2>&1 | grep -E "^n=" | sed 's/$/ (no refinement)/'
