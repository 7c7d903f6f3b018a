"""Prints the CTA-0 timeline of the pair-interaction kernel (SVGDB_TC_TRACE development aid)."""
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
_capi.load().svgdb_step(s._ctx, 2)
tr = np.loadtxt(os.environ["SVGDB_TC_TRACE"]).reshape(3, 64, 8)
t0 = tr[0, 0, 0]
names = ["mma", "wg0", "wg1"]
for t in range(4, 12):
    print("tile", t)
    print("  mma : pv0 wait+issue %6d  s0 %6d  pv1 %6d  s1 %6d   (start %d)" % (tr[0,t,1]-tr[0,t,0], tr[0,t,2]-tr[0,t,1], tr[0,t,3]-tr[0,t,2], tr[0,t,4]-tr[0,t,3], tr[0,t,0]-t0))
    for w in (1, 2):
        print("  %s : r_full wait %6d  s_full wait %6d  exp %6d  st_wait+arrive %6d  (start %d)" % (names[w], tr[w,t,1]-tr[w,t,0], tr[w,t,2]-tr[w,t,1], tr[w,t,3]-tr[w,t,2], tr[w,t,4]-tr[w,t,3], tr[w,t,0]-t0))
