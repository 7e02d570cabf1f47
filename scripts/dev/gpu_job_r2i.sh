#!/bin/bash
# round 2: parity suite with 2 and 4 K-blocks per fold in the accurate mode; ncu of the refinement block kernel
set -u
mkdir -p gpurun_out
V=$PWD/bayex_b200/csrc/build/variants
for c in 2 4; do
  BAYEX_B200_LIB=$V/acc_chj$c.so timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "production_path or explicit_point or config2 or config4 or non_default or fused or sample_matches or readme or 1d_optim or mixed_domain" > gpurun_out/r2i_pytest_chj$c.log 2>&1
  echo "pytest chj=$c rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2i_pytest_chj$c.log | tail -12
done
python scripts/dev/refine_only.py 50 > gpurun_out/r2i_refine_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:refine_blk -s 2 -c 2 -o gpurun_out/r2i_refine_blk python scripts/dev/refine_only.py 50 > gpurun_out/r2i_refine_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/r2i_refine_ncu.log
