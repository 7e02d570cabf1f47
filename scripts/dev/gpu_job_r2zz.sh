#!/bin/bash
# r2zz: single-launch refinement for n <= 256 - contract tests, every test that runs sample() at small n, sample() timeline
# with the single-launch kernel and (BX_NO_REFINE_SMALL=1) with the per-round path
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -k "refinement_contract or readme or reference_test or single_observation or optimize_suggestion or sample_matches_oracle or batched_proposals or state_file or state_rebuilds or ragged_shards or config2_sample or kept_hypers or fused_topk_mass" > gpurun_out/r2zz_tests.txt 2>&1
echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed|Error" gpurun_out/r2zz_tests.txt | tail -12
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python scripts/dev/small_sample_trace.py > gpurun_out/r2zz_small_sample_trace.txt 2>&1; grep -E "^==|wall" gpurun_out/r2zz_small_sample_trace.txt
echo "--- per-round path"
BX_NO_REFINE_SMALL=1 timeout 300 python scripts/dev/small_sample_trace.py > gpurun_out/r2zz_small_sample_trace_per_round.txt 2>&1; grep -E "^==|wall" gpurun_out/r2zz_small_sample_trace_per_round.txt
