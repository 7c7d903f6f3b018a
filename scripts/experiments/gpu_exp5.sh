#!/usr/bin/env bash
set -u
OUT=gpurun_out/exp5; mkdir -p $OUT
export SVGDB_PHI_F8=1
timeout 600 python -m pytest tests/test_gpu_tc32.py -q -s -p no:cacheprovider -k "phi_matches and 1-" --deselect "tests/test_gpu_tc32.py::test_tc32_phi_matches_oracle[1-513-2]" > $OUT/pytest.log 2>&1; echo "pytest exit $?" >> $OUT/pytest.log
grep -E "variant 1|passed|failed|exit|rror|timed out" $OUT/pytest.log | grep -v "print\|assert" | tail -14
timeout 300 python scripts/dbg_fullsize.py c3 1 128
SVGDB_PHI_F8=0 timeout 300 python scripts/dbg_fullsize.py c3 1 128
PYTEST=0 bash scripts/gpu_phi2.sh exp5_f8 "0 0"
SVGDB_PHI_F8=0 PYTEST=0 bash scripts/gpu_phi2.sh exp5_nof8 "0 0"
