#!/usr/bin/env bash
# One gpurun call: GPU parity tests, smoke, bench (N=1), ncu launch list and one full capture of the
# top kernel.  Everything lands in gpurun_out/.  Usage: gpurun --timeout 1500 -- bash scripts/gpu_check.sh [tag]
set -u
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,memory.total --format=csv > "$OUT/gpu.txt" 2>&1
python -c "import __graft_entry__ as g; g.build()" > "$OUT/build.log" 2>&1
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider --durations=10 > "$OUT/pytest_gpu.log" 2>&1
echo "pytest exit $?" >> "$OUT/pytest_gpu.log"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/smoke.log" 2>&1
echo "smoke exit $?" >> "$OUT/smoke.log"
timeout 600 python bench.py --steps 10 --warmup 3 > "$OUT/bench.json" 2> "$OUT/bench.err"
echo "bench exit $?" >> "$OUT/bench.err"
if [ "${SKIP_NCU:-0}" != "1" ]; then
  timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > "$OUT/ncu_plain.log" 2>&1 &&
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
      --log-file "$OUT/launches.csv" python bench.py --steps 2 --warmup 1 --no-cpu-baseline > "$OUT/ncu_launches.log" 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:phi_ -s 2 -c 1 \
      -o "$OUT/prof_phi" python bench.py --steps 2 --warmup 1 --no-cpu-baseline > "$OUT/ncu_full.log" 2>&1
fi
tail -5 "$OUT/pytest_gpu.log"; cat "$OUT/smoke.log" | tail -2; cat "$OUT/bench.json"; tail -3 "$OUT/bench.err"
