#!/usr/bin/env bash
# Quick GPU iteration: build, run the given pytest selection, optionally a bench line.
# Usage: gpurun --timeout 900 -- bash scripts/gpu_quick.sh <tag> "<pytest args>" [bench args]
set -u
TAG=${1:-quick}; shift
PYT=${1:-"tests -m gpu"}; shift || true
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
python -c "import __graft_entry__ as g; g.build()" > "$OUT/build.log" 2>&1 || cat "$OUT/build.log"
timeout ${PYTEST_TIMEOUT:-240} python -m pytest $PYT -q -s -p no:cacheprovider > "$OUT/pytest.log" 2>&1
echo "pytest exit $?" >> "$OUT/pytest.log"
grep -E "rel err|passed|failed|Error|error|exit" "$OUT/pytest.log" | tail -40
if [ $# -gt 0 ]; then
  timeout 600 python bench.py "$@" > "$OUT/bench.json" 2> "$OUT/bench.err"; echo "bench exit $?"; cat "$OUT/bench.json"; tail -5 "$OUT/bench.err"
fi
