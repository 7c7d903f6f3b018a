#!/usr/bin/env bash
# multi-GPU: parity check at NG ranks, then bench lines under torch.distributed.run.
#   usage: gpu_multi.sh NG "<workload> [ENV=VAL ...]" ...      (SKIP_CHECK=1 skips the parity check)
set -u
NG=${1:-2}; shift || true
OUT=gpurun_out/multi$NG; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
if [ "${SKIP_CHECK:-0}" != "1" ]; then
  for prec in ${CHECK_PRECS:-tc32 f64}; do
    timeout 600 $TR --master-port 29540 tests/multi_gpu_check.py --precision $prec > $OUT/check_$prec.log 2>&1; echo "check $prec exit $?" | tee -a $OUT/check_$prec.log
    grep -E "world=" $OUT/check_$prec.log | tail -8
  done
fi
k=0
for spec in "$@"; do
  set -- $spec; wl=$1; shift; k=$((k+1))
  tag=${wl}_$k
  env "$@" timeout 900 $TR --master-port $((29600+k)) bench.py --gpus $NG --workload $wl --steps 20 --warmup 10 --no-cpu-baseline > $OUT/bench_$tag.json 2> $OUT/bench_$tag.err
  echo "bench $wl $* exit $?"; grep -v "OMP_NUM_THREADS\|\*\*\*\*" $OUT/bench_$tag.err | tail -2 | cut -c1-300
  python - $OUT/bench_$tag.json <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith("{")][-1])
    print("  n_gpus %d ms/step %.3f e2e %.3f kernel_ms %.3f phases %s parity_ok %s phi_err %.3g passes/step %.2f" % (d["n_gpus"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["kernel_ms"], {k: round(v, 3) for k, v in d["roofline"]["phase_ms_per_step"].items()}, d["parity"]["ok"], d["parity"]["phi_sampled_rows_max_err_over_max_phi"], d["config"]["median_passes_per_step"]))
except Exception as e:
    print("  no bench line:", e)
PY
done
