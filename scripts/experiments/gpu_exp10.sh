#!/usr/bin/env bash
# FAST + e5m2 lo term, with / without the E . v_lo term: errors (small N forced, full size sampled) and speed
set -u
OUT=gpurun_out/exp10; mkdir -p $OUT
for nv in 1 0; do
  export SVGDB_PHI_NO_VLO=$nv
  echo "== SVGDB_PHI_NO_VLO=$nv"
  SVGDB_PHI_F8=1 timeout 600 python -m pytest tests/test_gpu_tc32.py -q -s -p no:cacheprovider -k "phi_matches and 1-" 2>&1 | grep -E "^\.?variant 1|passed|failed" | sed 's/(.*//'
  timeout 300 python scripts/dbg_fullsize.py c3 1 128
  timeout 300 python scripts/dbg_fullsize.py mvn48 1 128
  PYTEST=0 bash scripts/gpu_phi2.sh exp10_$nv "1 0" "1 1"
done
