#!/usr/bin/env bash
set -u
TAG=${1:-dist2}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
timeout 600 python -m pytest tests/test_gpu_tc32.py -q -x -s -p no:cacheprovider > "$OUT/pytest_tc32.log" 2>&1; echo "pytest exit $?" >> "$OUT/pytest_tc32.log"
grep -E "a rel err|passed|failed|exit|rror" "$OUT/pytest_tc32.log" | tail -24
for dk in 2 1; do
  SVGDB_DIST_KERNEL=$dk timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > "$OUT/bench_dk$dk.json" 2> "$OUT/bench_dk$dk.err"
  echo "dist kernel $dk: exit $?"; python - "$OUT/bench_dk$dk.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read())
    print("  ms/step %.3f  phases %s  passes/step %s finite %s" % (d["ms_per_step"], d["roofline"]["phase_ms_per_step"], d["config"]["median_passes_per_step"], d["config"]["finite"]))
except Exception as e:
    print("  no bench line:", e)
PY
  tail -2 "$OUT/bench_dk$dk.err"
done
