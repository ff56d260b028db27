#!/bin/bash
# compute-sanitizer (memcheck, then racecheck on shared memory) over the scoring and step-engine parity tests
set -u
mkdir -p gpurun_out
K="scoring_matches_reference_loop or score_pipeline or fused_update or engine_step or classif or f1 or philox"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider -k "$K" > gpurun_out/sanitize_memcheck.log 2>&1
echo "memcheck exit $?" >> gpurun_out/sanitize_memcheck.log
grep -E "ERROR SUMMARY|passed|failed|exit" gpurun_out/sanitize_memcheck.log | tail -5
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 3 python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider -k "scoring_matches_reference_loop and tc" > gpurun_out/sanitize_racecheck.log 2>&1
echo "racecheck exit $?" >> gpurun_out/sanitize_racecheck.log
grep -E "RACECHECK SUMMARY|hazard|passed|failed|exit" gpurun_out/sanitize_racecheck.log | sort | uniq -c | sort -rn | head -12
