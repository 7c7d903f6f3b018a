#!/usr/bin/env bash
# Late round 2: the pair kernel with the e5m2 correction term and the polynomial share (ncu --set full, after the plain run exited 0)
set -u
OUT=gpurun_out/r02prof2; mkdir -p $OUT
CMD="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-f64-leg --no-parity"
timeout 300 $CMD > $OUT/bench_plain.json 2> $OUT/bench_plain.err && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:phi2_tc32 -s 6 -c 1 -o $OUT/prof_phi2 $CMD > $OUT/ncu_phi2.log 2>&1
tail -3 $OUT/ncu_phi2.log; cat $OUT/bench_plain.json | head -c 600
