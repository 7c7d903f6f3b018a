#!/usr/bin/env bash
# timeline traces of CTA 0 for a list of SVGDB_PHI_DBG modes
set -u
TAG=${1:-trace}; shift
mkdir -p gpurun_out/$TAG
for dbg in "$@"; do
  SVGDB_PHI_DBG=$dbg SVGDB_TC_TRACE=gpurun_out/$TAG/trace_d$dbg.txt timeout 120 python scripts/tc_trace.py > gpurun_out/$TAG/trace_d$dbg.log 2>&1
  echo "== dbg $dbg"; python scripts/tc_trace_decode.py gpurun_out/$TAG/trace_d$dbg.txt 6 9
done
