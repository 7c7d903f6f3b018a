"""Decode a SVGDB_TC_TRACE file of phi2_tc32_kernel (CTA 0, i-tile 0's MMA warp and both exp warpgroups)."""
import sys
import numpy as np
tr = np.loadtxt(sys.argv[1]).reshape(3, 64, 8)
lo, hi = int(sys.argv[2]) if len(sys.argv) > 2 else 6, int(sys.argv[3]) if len(sys.argv) > 3 else 10
t0 = tr[0, lo, 0]
for t in range(lo, hi):
    m = tr[0, t] - t0
    print("tile %2d mma0: start %6d | unit k0 issued %6d | unit k1 issued %6d   (period %d)" % (t, m[0], m[1], m[2], tr[0, t + 1, 0] - tr[0, t, 0]))
    for w in (1, 2):
        e = tr[w, t] - t0
        print("     wg%d: k0 wait %6d got %6d arrived %6d | k1 wait %6d got %6d arrived %6d" % (w - 1, *e[1:7]))
