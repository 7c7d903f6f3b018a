#!/usr/bin/env bash
# One gpurun call: GPU parity tests, smoke, bench (N=1).  Everything lands in gpurun_out/<tag>/.
#   usage: gpurun --timeout 1500 -- bash scripts/gpu_r2_check.sh <tag> [pytest args...]
set -u
TAG=${1:-r02}; shift || true
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,memory.total --format=csv > "$OUT/gpu.txt" 2>&1
nproc >> "$OUT/gpu.txt"
python -c "import __graft_entry__ as g; g.build()" > "$OUT/build.log" 2>&1
if [ "${SKIP_PYTEST:-0}" != "1" ]; then
  timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider -s --durations=15 "$@" > "$OUT/pytest_gpu.log" 2>&1
  echo "pytest exit $?" >> "$OUT/pytest_gpu.log"
fi
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/smoke.log" 2>&1
echo "smoke exit $?" >> "$OUT/smoke.log"
if [ "${SKIP_BENCH:-0}" != "1" ]; then
  timeout 600 python bench.py --steps 20 --warmup 5 > "$OUT/bench.json" 2> "$OUT/bench.err"
  echo "bench exit $?" >> "$OUT/bench.err"
fi
grep -E "passed|failed|error|exit" "$OUT/pytest_gpu.log" | tail -8; tail -3 "$OUT/smoke.log"; cut -c1-1500 "$OUT/bench.json"; tail -3 "$OUT/bench.err"
