#!/bin/bash
# compute-sanitizer memcheck over the parity tests of the step path (pair kernels, fused update incl. the in-process
# multi-rank exchanges, FeatNet kernels, sampler, host pipe) and the scoring path; racecheck on the tensor-core scoring.
set -u
O=gpurun_out/${TAG:-r2san}; mkdir -p $O
K="update_rows or multi_rank or exchange or featnet or philox or host_pipe or engine_step or fused_update or pipelined or wordnet"
timeout ${MEMCHECK_S:-420} compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider -k "$K" > $O/sanitize_memcheck_step.log 2>&1
echo "memcheck(step) exit $?" >> $O/sanitize_memcheck_step.log
grep -E "ERROR SUMMARY|passed|failed|exit" $O/sanitize_memcheck_step.log | tail -4
if [ "${SCORING:-1}" = "1" ]; then
  timeout ${MEMCHECK_S:-420} compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider -k "scoring_matches_reference_loop or score_pipeline or classif or f1" > $O/sanitize_memcheck_score.log 2>&1
  echo "memcheck(scoring) exit $?" >> $O/sanitize_memcheck_score.log
  grep -E "ERROR SUMMARY|passed|failed|exit" $O/sanitize_memcheck_score.log | tail -4
fi
