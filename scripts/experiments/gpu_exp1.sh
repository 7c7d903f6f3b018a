#!/usr/bin/env bash
# experiment: folded distance pass (correctness + timing) and pair-kernel variants
set -u
OUT=gpurun_out/exp1; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_tc32.py -q -x -s -p no:cacheprovider -k "phi_matches or narrowing or full_size_properties or chunked or step_host" > $OUT/pytest.log 2>&1; echo "pytest exit $?" >> $OUT/pytest.log
grep -E "passed|failed|exit|rror" $OUT/pytest.log | tail -5
for fold in 1 0; do
  echo "== SVGDB_DIST_FOLD=$fold"; SVGDB_DIST_FOLD=$fold timeout 300 python scripts/gpu_time_kernels.py 2>&1 | tee $OUT/time_fold$fold.log
done
PYTEST=0 bash scripts/gpu_phi2.sh exp1 "0 0" "0 1" "0 2" "4 0" "8 0"
