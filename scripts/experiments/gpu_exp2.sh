#!/usr/bin/env bash
set -u
OUT=gpurun_out/exp2; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_tc32.py -q -x -s -p no:cacheprovider > $OUT/pytest.log 2>&1; echo "pytest exit $?" >> $OUT/pytest.log
grep -E "passed|failed|exit|rror|a rel err|N=65536" $OUT/pytest.log | grep -v "print\|assert" | tail -30
for f in 1 0; do echo "== SVGDB_DIST_F16=$f"; SVGDB_DIST_F16=$f timeout 300 python scripts/gpu_time_kernels.py 2>&1 | grep -E "variant 0|pair" ; done
PYTEST=0 bash scripts/gpu_phi2.sh exp2 "0 0"
