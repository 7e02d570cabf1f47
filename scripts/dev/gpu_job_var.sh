#!/bin/bash
# kernel time of the bench configuration for the base library and the variants given as arguments
set -u
V=bayex_b200/csrc/build/variants
for name in base "$@"; do
  if [ "$name" == "base" ]; then unset BAYEX_B200_LIB; else export BAYEX_B200_LIB=$PWD/$V/$name.so; fi
  timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$name kernel_ms', round(d['roofline']['kernel_ms'],1), 'ms/step', round(d['ms_per_step'],1), d['clocks']['sm_mhz'])"
done
