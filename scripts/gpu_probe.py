"""One small TC32 ComputePhi; prints the error text if a launch fails (development aid)."""
import sys
import numpy as np
sys.path.insert(0, ".")
import svgdcpp_b200 as sv
n, d = 300, 64
rng = np.random.default_rng(0)
x0 = np.asfortranarray(rng.standard_normal((d, n)))
model = sv.MultivariateNormal(np.zeros(d), np.eye(d))
s = sv.SVGD(d, 1, x0, sv.GaussianRBFKernel(x0, sv.ScaleMethod.Median, model), model, sv.Adam(d, n, 0.1, 0.9, 0.999), precision=1)
try:
    phi, a = s.ComputePhi()
    print("ok", a, float(np.abs(phi).max()))
except Exception as e:
    print("FAILED:", e)
