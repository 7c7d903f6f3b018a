#!/usr/bin/env bash
# Round-2 evidence run on one B200: bench line (not under a profiler), ncu launch list of the steady state, full captures of
# the three hot kernels (d <= 64 pair kernel, distance pass; wide pair kernel on a config-4 slice).
set -u
OUT=gpurun_out/r02prof; mkdir -p $OUT
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt 2>&1
CMD="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-f64-leg --no-parity"
timeout 300 $CMD > $OUT/bench_plain.json 2> $OUT/bench_plain.err && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 260 --csv --log-file $OUT/launches_steady.csv $CMD > $OUT/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:phi2_tc32 -s 6 -c 1 -o $OUT/prof_phi2 $CMD > $OUT/ncu_phi2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dist2_tc32 -s 12 -c 1 -o $OUT/prof_dist2 $CMD > $OUT/ncu_dist2.log 2>&1
CMDW="python bench.py --workload c4 --particles 32768 --steps 6 --warmup 8 --no-cpu-baseline --no-parity"
timeout 300 $CMDW > $OUT/bench_c4slice_plain.json 2> $OUT/bench_c4slice_plain.err && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:phiw_tc32 -s 10 -c 1 -o $OUT/prof_phiw $CMDW > $OUT/ncu_phiw.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:distw_tc32 -s 12 -c 1 -o $OUT/prof_distw $CMDW > $OUT/ncu_distw.log 2>&1
ls -la $OUT | head -30
