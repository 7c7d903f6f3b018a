#!/usr/bin/env bash
# mixture gradient through library DGEMMs: parity tests, then config 4 with and without
set -u
OUT=gpurun_out/exp17; mkdir -p $OUT
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_wide.py -q -s -p no:cacheprovider -k "mixture or gmm or wide or config4 or slice" > $OUT/pytest.log 2>&1; echo "pytest exit $?" >> $OUT/pytest.log
grep -E "gemm form|passed|failed|exit|rror" $OUT/pytest.log | cut -c1-200 | tail -12
timeout 300 python bench.py --workload c4 --steps 8 --warmup 8 --no-cpu-baseline > $OUT/bench_c4_gemm.json 2> $OUT/bench_c4_gemm.err; echo "bench c4 (gemm) exit $?"
SVGDB_GRAD_GEMM=0 timeout 300 python bench.py --workload c4 --steps 8 --warmup 8 --no-cpu-baseline --no-parity > $OUT/bench_c4_old.json 2> $OUT/bench_c4_old.err; echo "bench c4 (one-kernel) exit $?"
python - <<'PY'
import json
for f in ("bench_c4_gemm", "bench_c4_old"):
    try:
        d = json.loads(open("gpurun_out/exp17/%s.json" % f).read())
        print(f, "ms/step %.1f e2e %.1f kernel %.1f phases %s parity %s" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["kernel_ms"], {k: round(v, 1) for k, v in d["roofline"]["phase_ms_per_step"].items()}, d["parity"]))
    except Exception as e:
        print(f, "no line", e)
PY
