#!/usr/bin/env bash
# one-GPU evidence: every GPU test, smoke, the bench lines of both workloads, the C5 sweep column
set -u
OUT=gpurun_out/final1; mkdir -p $OUT
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt 2>&1; nproc >> $OUT/gpu.txt
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider -s --durations=10 > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
grep -E "passed|failed|exit" $OUT/pytest_gpu.log | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke exit $?" >> $OUT/smoke.log; tail -3 $OUT/smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/bench_c3.json 2> $OUT/bench_c3.err; echo "bench c3 exit $?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "bench ref exit $?"
timeout 900 python bench.py --workload c4 --steps 10 --warmup 10 > $OUT/bench_c4.json 2> $OUT/bench_c4.err; echo "bench c4 exit $?"
timeout 1200 python scripts/sweep.py --cpu --budget 25 > $OUT/sweep_1gpu.jsonl 2> $OUT/sweep.err; echo "sweep exit $?"; wc -l $OUT/sweep_1gpu.jsonl
python - <<'PY'
import json
for f in ("bench_c3", "bench_c4"):
    try:
        d = json.loads(open("gpurun_out/final1/%s.json" % f).read())
        print(f, "ms/step %.3f e2e %.3f kernel %.3f frac_burst %.3f step_frac_burst %.3f parity %s cpu %s" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac_burst"], d["roofline"]["whole_step_frac_burst"], d["parity"]["ok"], d.get("cpu_baseline", {}).get("value")))
    except Exception as e:
        print(f, "no line", e)
PY
