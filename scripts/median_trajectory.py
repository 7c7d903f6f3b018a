"""Records the median of D2 per step at the bench shape (to study bracket predictors offline)."""
import json, sys
import numpy as np
sys.path.insert(0, '.')
import svgdcpp_b200 as sv
from svgdcpp_b200 import synth, _capi
n, d = 65536, 64
x0, means, covs = synth.mvn_problem(n, d)
model = sv.MultivariateNormal(means[0], covs[0])
s = sv.SVGD(d, 1, x0, sv.GaussianRBFKernel(x0, sv.ScaleMethod.Median, model), model, sv.Adam(d, n, 0.1, 0.9, 0.999), precision=1)
s.Initialize()
lib = _capi.load(); s._upload()
out = []
for it in range(120):
    lib.svgdb_step(s._ctx, 1)
    st = s.Stats()
    out.append({"it": it, "med2": float(np.log(n) / st['last_scale']), "passes": int(st['median_passes']), "hits": int(st['median_bracket_hits'])})
json.dump(out, open("gpurun_out/med_traj.json", "w"))
print("done", out[-1])
