#!/usr/bin/env bash
# multi-GPU: parity tests at every available rank count, then bench lines (c3 and c4) under torch.distributed.run
set -u
NG=${1:-2}; shift || true
OUT=gpurun_out/multi$NG; mkdir -p $OUT
if [ "${SKIP_PYTEST:-0}" != "1" ]; then
  timeout 900 python -m pytest tests/test_gpu_multi.py -q -x -s -p no:cacheprovider > $OUT/pytest.log 2>&1; echo "pytest exit $?" >> $OUT/pytest.log
  grep -E "world=|passed|failed|exit|rror" $OUT/pytest.log | tail -40
fi
for wl in "$@"; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29600 bench.py --gpus $NG --workload $wl --steps 20 --warmup 10 --no-cpu-baseline > $OUT/bench_$wl.json 2> $OUT/bench_$wl.err
  echo "bench $wl exit $?"; tail -2 $OUT/bench_$wl.err | cut -c1-300
  python - $OUT/bench_$wl.json <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith("{")][-1])
    print("  n_gpus %d ms/step %.3f e2e %.3f kernel_ms %.3f phases %s parity %s" % (d["n_gpus"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["kernel_ms"], {k: round(v, 3) for k, v in d["roofline"]["phase_ms_per_step"].items()}, d["parity"]))
except Exception as e:
    print("  no bench line:", e)
PY
done
