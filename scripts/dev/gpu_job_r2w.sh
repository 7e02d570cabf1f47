#!/bin/bash
# r2w: expanded squared distance in the fast mode's K* producers: A/B (BX_EXPAND_SMAX=0 = direct form) + full GPU suite
mkdir -p gpurun_out
echo "== BX_EXPAND_SMAX=0 (direct form on centred rows)" > gpurun_out/r2w_probe.txt
BX_EXPAND_SMAX=0 timeout 300 python scripts/dev/probe_expand.py >> gpurun_out/r2w_probe.txt 2>&1
echo "== default (expanded form while |c'|^2 + |u'|^2 <= 64)" >> gpurun_out/r2w_probe.txt
timeout 300 python scripts/dev/probe_expand.py >> gpurun_out/r2w_probe.txt 2>&1
cat gpurun_out/r2w_probe.txt
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2w_tests.txt 2>&1
tail -8 gpurun_out/r2w_tests.txt
