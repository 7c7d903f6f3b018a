#!/usr/bin/env bash
# e5m2 correction terms: sampled phi rows (incl. the outermost particle) at N = 65536 for several dimensions, forced on / off / automatic
set -u
for wl in mvn2 mvn8 mvn16 mvn32 c3; do
  for f8 in 1 0; do
    echo "SVGDB_PHI_F8=$f8"; SVGDB_PHI_F8=$f8 timeout 300 python scripts/dbg_fullsize.py $wl 1 128
  done
done
echo automatic; timeout 300 python scripts/dbg_fullsize.py c3 1 128
timeout 900 python -m pytest tests/test_gpu_tc32.py -q -x -p no:cacheprovider 2>&1 | tail -3
