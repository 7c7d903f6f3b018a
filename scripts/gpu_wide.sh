#!/usr/bin/env bash
set -u
OUT=gpurun_out/wide; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_wide.py -q -s -p no:cacheprovider "$@" > $OUT/pytest.log 2>&1; echo "pytest exit $?" >> $OUT/pytest.log
grep -E "wide|C4|passed|failed|exit|rror|timed out" $OUT/pytest.log | grep -v "^    \|print(" | head -60
