#!/usr/bin/env bash
set -u
OUT=gpurun_out/exp4; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_tc32.py -q -x -p no:cacheprovider > $OUT/pytest.log 2>&1; echo "pytest exit $?" >> $OUT/pytest.log
grep -E "passed|failed|exit|rror|timed out" $OUT/pytest.log | tail -5
for cl in 1 0; do
  echo "== SVGDB_PHI_CLUSTER=$cl"; SVGDB_PHI_CLUSTER=$cl PYTEST=0 bash scripts/gpu_phi2.sh exp4_cl$cl "0 0" "0 2"
  SVGDB_PHI_CLUSTER=$cl SVGDB_TC32_VARIANT=2 PYTEST=0 bash scripts/gpu_phi2.sh exp4_cl${cl}_precise "0 0"
done
