"""Kernel-only timings at the bench shape through svgdb_time_kernel (development / profiling aid)."""
import ctypes as C
import sys
sys.path.insert(0, ".")
import svgdcpp_b200 as sv
from svgdcpp_b200 import synth, _capi
n, d = 65536, 64
x0, means, covs = synth.mvn_problem(n, d)
model = sv.MultivariateNormal(means[0], covs[0])
s = sv.SVGD(d, 1, x0, sv.GaussianRBFKernel(x0, sv.ScaleMethod.Median, model), model, sv.Adam(d, n, 0.1, 0.9, 0.999), precision=1)
s.Initialize(); s._upload()
lib = _capi.load()
assert lib.svgdb_step(s._ctx, 6) == 0
ms = C.c_float()
for rel in (7.8e-4, 1e-4, 2e-5):
    for variant in (0, 2, 1, 3):
        rc = lib.svgdb_time_kernel(s._ctx, 0, 5, variant, rel, C.byref(ms))
        print("dist pass  rel half-width %.1e variant %d: %.3f ms (rc %d)" % (rel, variant, ms.value, rc))
rc = lib.svgdb_time_kernel(s._ctx, 1, 5, 0, 0.0, C.byref(ms))
print("pair pass (prep + kernel + epilogue kernel): %.3f ms (rc %d)" % (ms.value, rc))
