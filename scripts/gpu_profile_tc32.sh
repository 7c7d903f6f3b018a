#!/usr/bin/env bash
# TC32 evidence run: GPU tests, bench line, ncu launch list and full captures of the two tensor-core kernels.
set -u
TAG=${1:-r01_tc32}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
python -c "import __graft_entry__ as g; g.build()" > "$OUT/build.log" 2>&1
timeout 600 python -m pytest tests -q -m gpu -p no:cacheprovider > "$OUT/pytest_gpu.log" 2>&1; echo "pytest exit $?" >> "$OUT/pytest_gpu.log"
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/smoke.log" 2>&1; echo "smoke exit $?" >> "$OUT/smoke.log"
timeout 400 python bench.py --steps 20 --warmup 5 > "$OUT/bench.json" 2> "$OUT/bench.err"; echo "bench exit $?" >> "$OUT/bench.err"
timeout 400 python bench.py --steps 10 --warmup 3 --precision f64 --no-cpu-baseline > "$OUT/bench_f64.json" 2>> "$OUT/bench.err"
timeout 200 python scripts/gpu_time_kernels.py > "$OUT/kernel_times.log" 2>&1
timeout 200 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > "$OUT/ncu_plain.log" 2>&1 &&
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file "$OUT/launches.csv" python bench.py --steps 4 --warmup 3 --no-cpu-baseline > "$OUT/ncu_launches.log" 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"phi2_tc32" -s 3 -c 1 -o "$OUT/prof_phi_tc32" python bench.py --steps 4 --warmup 3 --no-cpu-baseline > "$OUT/ncu_full_phi.log" 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"dist2_tc32" -s 10 -c 1 -o "$OUT/prof_dist_tc32" python bench.py --steps 4 --warmup 3 --no-cpu-baseline > "$OUT/ncu_full_dist.log" 2>&1
tail -4 "$OUT/pytest_gpu.log"; tail -2 "$OUT/smoke.log"; cat "$OUT/bench.json"; tail -2 "$OUT/bench.err"; cat "$OUT/kernel_times.log"
