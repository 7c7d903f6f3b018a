#!/usr/bin/env bash
set -u
for wl in mvn64n16384 mvn64n32768 mvn48n16384 mvn64n131072; do
  for nv in 1 0; do
    echo -n "NO_VLO=$nv  "; SVGDB_PHI_F8=1 SVGDB_PHI_NO_VLO=$nv timeout 300 python scripts/dbg_fullsize.py $wl 1 128
  done
done
