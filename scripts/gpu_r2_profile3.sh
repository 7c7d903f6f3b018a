#!/usr/bin/env bash
# steady-state (gated, folded, fp16) distance pass under ncu --set full, after the plain run exited 0
set -u
OUT=gpurun_out/r02prof3; mkdir -p $OUT
CMD="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-f64-leg --no-parity"
timeout 300 $CMD > $OUT/bench_plain.json 2> $OUT/bench_plain.err && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dist2_tc32 -s 30 -c 1 -o $OUT/prof_dist2 $CMD > $OUT/ncu_dist2.log 2>&1
tail -2 $OUT/ncu_dist2.log
