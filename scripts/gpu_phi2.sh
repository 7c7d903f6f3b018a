#!/usr/bin/env bash
# A/B run of the pair-interaction kernel variants: short bench lines per variant.
#   usage: gpu_phi2.sh <tag> "<poly dbg>" ...   (SVGDB_PHI_POLY: exponential pairs per chunk on the FMA pipe; SVGDB_PHI_DBG: see Phi2Args::dbg)
set -u
TAG=${1:-phi2}; shift
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
if [ "${PYTEST:-1}" = "1" ]; then
  timeout 600 python -m pytest tests/test_gpu_tc32.py -q -x -s -p no:cacheprovider > "$OUT/pytest_tc32.log" 2>&1; echo "pytest exit $?" >> "$OUT/pytest_tc32.log"
  grep -E "phi max-rel|trajectory|passed|failed|exit|rror" "$OUT/pytest_tc32.log" | tail -20
fi
for cfg in "$@"; do
  set -- $cfg
  F="$OUT/bench_p$1_d$2"
  SVGDB_PHI_POLY=$1 SVGDB_PHI_DBG=$2 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-f64-leg --no-parity > "$F.json" 2> "$F.err"
  echo "poly $1 dbg $2: exit $?"; python - "$F.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read())
    print("  ms/step %.3f  phi %.3f ms  median %.3f ms  frac %.3f  finite %s" % (d["ms_per_step"], d["roofline"]["phase_ms_per_step"]["phi"], d["roofline"]["phase_ms_per_step"]["median"], d["roofline"]["frac_sustained"], d["config"]["finite"]))
except Exception as e:
    print("  no bench line:", e)
PY
  tail -2 "$F.err"
done
if [ "${TRACE:-0}" = "1" ]; then
  SVGDB_TC_TRACE=gpurun_out/$TAG/trace.txt timeout 120 python scripts/tc_trace.py > "$OUT/trace.log" 2>&1; tail -3 "$OUT/trace.log"
fi
