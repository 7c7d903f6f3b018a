#!/usr/bin/env bash
set -u
OUT=gpurun_out/exp3; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_wide.py -q -x -s -p no:cacheprovider > $OUT/pytest.log 2>&1; echo "pytest exit $?" >> $OUT/pytest.log
grep -E "passed|failed|exit|rror|timed out" $OUT/pytest.log | tail -5
for cl in 1 0; do
  echo "== SVGDB_PHI_CLUSTER=$cl"
  SVGDB_PHI_CLUSTER=$cl timeout 300 python scripts/dbg_fullsize.py c4s 2 64
  SVGDB_PHI_CLUSTER=$cl timeout 300 python scripts/dbg_fullsize.py c4s 1 128
done
