#!/usr/bin/env bash
set -u
TAG=${1:-ncu_dist}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
timeout 200 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > "$OUT/plain.log" 2>&1 || { tail -5 "$OUT/plain.log"; exit 1; }
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/launches.csv" python bench.py --steps 4 --warmup 3 --no-cpu-baseline > "$OUT/ncu_launches.log" 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"dist2_tc32" -s 10 -c 1 -o "$OUT/prof_dist2" python bench.py --steps 4 --warmup 3 --no-cpu-baseline > "$OUT/ncu_full.log" 2>&1
tail -2 "$OUT/ncu_full.log"; ls -la "$OUT"
