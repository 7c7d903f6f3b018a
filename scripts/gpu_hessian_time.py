"""Step time of the Hessian-scaled kernel (ScaleMethod::Hessian) at the headline shape, both precisions, next to the median scale."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import svgdcpp_b200 as sv
from svgdcpp_b200 import synth

n, d = 65536, 64
x0, means, covs = synth.mvn_problem(n, d)
model = sv.MultivariateNormal(means[0], covs[0])
for prec, name, steps in [(1, "tc32", 20), (0, "f64", 3)]:
    for method in (sv.ScaleMethod.Hessian, sv.ScaleMethod.Median):
        x = x0.copy(order="F")
        svgd = sv.SVGD(d, 1, x, sv.GaussianRBFKernel(x, method, model), model, sv.Adam(d, n, 0.1, 0.9, 0.999), precision=prec)
        svgd.Initialize()
        svgd.Step(3)
        t0 = time.perf_counter()
        svgd.Step(steps)
        dt = (time.perf_counter() - t0) / steps
        svgd._lib.svgdb_set_profiling(svgd._ctx, 1)
        before = svgd.Stats()
        svgd.Step(steps)
        after = svgd.Stats()
        phases = {k: round((after[k] - before[k]) / steps, 3) for k in after if k.startswith("ms_")}
        print("   phases (CUDA events; for Hessian ms_median = Hessian sum + Cholesky, ms_comm = change of variables + operands):", phases)
        print("%s %s: %.3f ms/step (host clock, particles read back once per Step call), finite %s" % (name, method.name, dt * 1e3, bool(np.all(np.isfinite(x)))), flush=True)
        svgd.close()

# a sum of three overlapping Gaussians: the Hessian sum kernel runs every step
rng = np.random.default_rng(1)
C = 3
mus = 0.4 * rng.standard_normal((C, d))
cvs = [(lambda M: M @ M.T / d + 0.7 * np.eye(d))(rng.standard_normal((d, d))) for _ in range(C)]
model = None
for k in range(C):
    m = sv.MultivariateNormal(mus[k], cvs[k])
    model = m if model is None else model + m
x = np.asfortranarray(1.2 * rng.standard_normal((d, n)))
svgd = sv.SVGD(d, 1, x, sv.GaussianRBFKernel(x, sv.ScaleMethod.Hessian, model), model, sv.Adam(d, n, 0.05, 0.9, 0.999), precision=1)
svgd.Initialize()
svgd.Step(3)
svgd._lib.svgdb_set_profiling(svgd._ctx, 1)
before = svgd.Stats()
svgd.Step(10)
after = svgd.Stats()
print("tc32 Hessian, 3 components:", {k: round((after[k] - before[k]) / 10, 3) for k in after if k.startswith("ms_")}, "finite", bool(np.all(np.isfinite(x))))
svgd.close()
