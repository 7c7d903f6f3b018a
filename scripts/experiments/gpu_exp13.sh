#!/usr/bin/env bash
set -u
for cfg in "0 0 0" "1 1 0" "1 1 1" "1 1 2"; do
  set -- $cfg
  SVGDB_PHI_F8=$1 SVGDB_PHI_NO_VLO=$2 SVGDB_PHI_DBG=$3 timeout 200 python scripts/tc_trace.py 2>&1 | tail -4
done
