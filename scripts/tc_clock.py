"""SM clock inside the pair kernel at the headline shape: cycles / nanoseconds over CTA 0's lifetime (SVGDB_TC_TRACE development aid)."""
import os, sys
import numpy as np
sys.path.insert(0, ".")
os.environ.setdefault("SVGDB_TC_TRACE", "gpurun_out/tc_trace.txt")
import svgdcpp_b200 as sv
from svgdcpp_b200 import synth, _capi
n, d = 65536, int(sys.argv[1]) if len(sys.argv) > 1 else 64
x0, means, covs = synth.mvn_problem(n, d)
model = sv.MultivariateNormal(means[0], covs[0])
s = sv.SVGD(d, 1, x0, sv.GaussianRBFKernel(x0, sv.ScaleMethod.Median, model), model, sv.Adam(d, n, 0.1, 0.9, 0.999), precision=1)
s.Initialize(); s._upload()
_capi.load().svgdb_step(s._ctx, 12)
tr = np.loadtxt(os.environ["SVGDB_TC_TRACE"]).reshape(3, 64, 8)
cyc = tr[1, 63, 7] - tr[1, 63, 0]; ns = tr[2, 63, 7] - tr[2, 63, 0]
units = (n / 256) * (n / 128) * 2 / 148
print("F8=%s NO_VLO=%s DBG=%s: CTA 0 lived %.0f cycles = %.3f ms -> %.3f GHz; %.0f cycles per unit" % (
    os.environ.get("SVGDB_PHI_F8", "-"), os.environ.get("SVGDB_PHI_NO_VLO", "-"), os.environ.get("SVGDB_PHI_DBG", "0"), cyc, ns * 1e-6, cyc / ns, cyc / units))
