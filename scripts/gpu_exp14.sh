#!/usr/bin/env bash
set -u
echo "== with row sum"; PYTEST=0 bash scripts/gpu_phi2.sh exp14a "0 0" "1 0" "2 0" "3 0"
echo "== without row sum (timing only)"; SVGDB_PHI_NOSUM=1 PYTEST=0 bash scripts/gpu_phi2.sh exp14b "0 0" "1 0" "2 0" "3 0" "4 0"
