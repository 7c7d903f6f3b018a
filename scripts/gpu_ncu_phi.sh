#!/usr/bin/env bash
# One ncu --set full capture (with source counters) of the pair-interaction kernel at the bench shape.
set -u
TAG=${1:-ncu_phi}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
timeout 200 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > "$OUT/plain.log" 2>&1 || { tail -5 "$OUT/plain.log"; exit 1; }
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"phi2_tc32" -s 3 -c 1 -o "$OUT/prof_phi2" python bench.py --steps 4 --warmup 3 --no-cpu-baseline > "$OUT/ncu.log" 2>&1
tail -3 "$OUT/ncu.log"; ls -la "$OUT"
