#!/usr/bin/env bash
set -u
for ts in 1 0; do
  export SVGDB_PHI_TCSUM=$ts
  echo "== SVGDB_PHI_TCSUM=$ts"
  SVGDB_PHI_F8=1 SVGDB_PHI_NO_VLO=0 timeout 600 python -m pytest tests/test_gpu_tc32.py -q -s -p no:cacheprovider -k "phi_matches and 1-" 2>&1 | grep -E "^\.?F?variant 1|passed|failed" | sed 's/(.*//'
  timeout 300 python scripts/dbg_fullsize.py c3 1 128
  timeout 300 python scripts/dbg_fullsize.py mvn64n16384 1 128
  PYTEST=0 bash scripts/gpu_phi2.sh exp15_$ts "1 0" "2 0" "3 0"
done
