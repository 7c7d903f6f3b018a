#!/usr/bin/env bash
# one-GPU evidence after the late-round pair-kernel changes: every GPU test, smoke, bench lines, launch list + full capture of the pair kernel
set -u
OUT=gpurun_out/final2; mkdir -p $OUT
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt 2>&1; nproc >> $OUT/gpu.txt
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider -s --durations=10 > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
grep -E "passed|failed|exit" $OUT/pytest_gpu.log | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke exit $?" >> $OUT/smoke.log; tail -3 $OUT/smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/bench_c3.json 2> $OUT/bench_c3.err; echo "bench c3 exit $?"
CMD="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-f64-leg --no-parity"
timeout 300 $CMD > $OUT/bench_c3_for_ncu.json 2> $OUT/bench_c3_for_ncu.err && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 260 --csv --log-file $OUT/launches_steady.csv $CMD > $OUT/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:phi2_tc32 -s 6 -c 1 -o $OUT/prof_phi2 $CMD > $OUT/ncu_phi2.log 2>&1
timeout 600 python scripts/sweep.py --no-f64 --ds 8,32,64 --budget 25 > $OUT/sweep_1gpu_tc32.jsonl 2> $OUT/sweep.err; echo "sweep exit $?"; wc -l $OUT/sweep_1gpu_tc32.jsonl
python - <<'PY'
import json
d = json.loads(open("gpurun_out/final2/bench_c3.json").read())
print("ms/step %.3f e2e %.3f kernel %.3f frac_burst %.3f parity %s cpu %s" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac_burst"], d["parity"], d.get("cpu_baseline", {}).get("value")))
PY
