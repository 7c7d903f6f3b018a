"""Prints the CTA-0 timeline of the pair-interaction kernel (SVGDB_TC_TRACE development aid): per j-tile, what the MMA issuer of
i-tile 0 and one exp warp of each i-tile spent waiting and working (cycles)."""
import os, sys
import numpy as np
sys.path.insert(0, ".")
os.environ.setdefault("SVGDB_TC_TRACE", "gpurun_out/tc_trace.txt")
import svgdcpp_b200 as sv
from svgdcpp_b200 import synth, _capi
n, d = 65536, 64
x0, means, covs = synth.mvn_problem(n, d)
model = sv.MultivariateNormal(means[0], covs[0])
s = sv.SVGD(d, 1, x0, sv.GaussianRBFKernel(x0, sv.ScaleMethod.Median, model), model, sv.Adam(d, n, 0.1, 0.9, 0.999), precision=1)
s.Initialize(); s._upload()
_capi.load().svgdb_step(s._ctx, 12)
tr = np.loadtxt(os.environ["SVGDB_TC_TRACE"]).reshape(3, 64, 8)
print("F8=%s NO_VLO=%s DBG=%s" % (os.environ.get("SVGDB_PHI_F8", "-"), os.environ.get("SVGDB_PHI_NO_VLO", "-"), os.environ.get("SVGDB_PHI_DBG", "0")))
m = tr[0]
ts = range(8, 56)
w0 = np.mean([m[t, 3] - m[t, 0] for t in ts]); i0 = np.mean([m[t, 1] - m[t, 3] for t in ts])
w1 = np.mean([m[t, 4] - m[t, 1] for t in ts]); i1 = np.mean([m[t, 2] - m[t, 4] for t in ts])
per = np.mean([m[t + 1, 0] - m[t, 0] for t in ts])
print("  MMA issuer of tile 0, per j-tile (2 units): %.0f cycles = wait E(b0) %.0f + issue PV,S' %.0f + wait E(b1) %.0f + issue %.0f" % (per, w0, i0, w1, i1))
for w in (1, 2):
    e = tr[w]
    # ev 2 + 3k: S_k in registers; 3 + 3k: E_k stored and handed on; 1 + 3 kn: started to wait for the next S (only if it was not there yet)
    ex0 = np.mean([e[t, 3] - e[t, 2] for t in ts]); ex1 = np.mean([e[t, 6] - e[t, 5] for t in ts])
    g0 = np.mean([e[t, 5] - e[t, 3] for t in ts]); g1 = np.mean([e[t + 1, 2] - e[t, 6] for t in ts])
    print("  exp warp of tile %d: unit 0 exp+store %.0f, then %.0f until S(1) is in registers, unit 1 %.0f, then %.0f until the next S(0)" % (w - 1, ex0, g0, ex1, g1))
