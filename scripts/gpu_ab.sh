#!/usr/bin/env bash
# A/B of two builds of the library in one call: lib/libsvgd_b200.so (B, current sources) against lib/libsvgd_b200_A.so (saved before the change)
set -u
L=svgdcpp_b200/lib
cp $L/libsvgd_b200.so $L/libsvgd_b200_B.so
for rep in 1 2; do
  for v in B A; do
    cp $L/libsvgd_b200_$v.so $L/libsvgd_b200.so
    echo "== build $v (rep $rep)"; PYTEST=0 bash scripts/gpu_phi2.sh ab_${v}_$rep "${1:-1 0}"
  done
done
cp $L/libsvgd_b200_B.so $L/libsvgd_b200.so
timeout 300 python scripts/dbg_fullsize.py c3 1 128
