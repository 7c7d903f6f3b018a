"""CTA-0 timeline of the tensor-core distance pass (SVGDB_TC_TRACE + SVGDB_TC_TRACE_DIST development aid)."""
import os, sys
import numpy as np
sys.path.insert(0, ".")
os.environ["SVGDB_TC_TRACE"] = "gpurun_out/tc_trace_dist.txt"
os.environ["SVGDB_TC_TRACE_DIST"] = "1"
import svgdcpp_b200 as sv
from svgdcpp_b200 import synth, _capi
n, d = 65536, 64
x0, means, covs = synth.mvn_problem(n, d)
model = sv.MultivariateNormal(means[0], covs[0])
s = sv.SVGD(d, 1, x0, sv.GaussianRBFKernel(x0, sv.ScaleMethod.Median, model), model, sv.Adam(d, n, 0.1, 0.9, 0.999), precision=1)
s.Initialize(); s._upload()
_capi.load().svgdb_step(s._ctx, 6)
tr = np.loadtxt("gpurun_out/tc_trace_dist.txt").reshape(3, 64, 8)
t0 = tr[0, 4, 0]
for t in range(4, 12):
    print("tile %2d mma: w0 wait_free %5d issue %5d | w1 wait_free %5d issue %5d (at %7d)   wg0: wait_s %5d count %5d (at %7d)  wg1: wait_s %5d count %5d" % (
        t, tr[0,t,1]-tr[0,t,0], tr[0,t,2]-tr[0,t,1], tr[0,t,4]-tr[0,t,3], tr[0,t,5]-tr[0,t,4], tr[0,t,0]-t0,
        tr[1,t,1]-tr[1,t,0], tr[1,t,2]-tr[1,t,1], tr[1,t,0]-t0, tr[2,t,1]-tr[2,t,0], tr[2,t,2]-tr[2,t,1]))
