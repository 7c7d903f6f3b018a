"""Whole-step time (svgdb_time_steps, CUDA events) over consecutive blocks of steps, with the median statistics per block."""
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
assert lib.svgdb_step(s._ctx, 5) == 0
ms = C.c_float()
prev = s.Stats()
for rep in range(8):
    assert lib.svgdb_time_steps(s._ctx, 20, C.byref(ms)) == 0
    st = s.Stats()
    print("block %d: %.3f ms/step, median passes %d, bracket hits %d, scale %.6g" % (
        rep, ms.value / 20, st["median_passes"] - prev["median_passes"], st["median_bracket_hits"] - prev["median_bracket_hits"], st["last_scale"]))
    prev = st
lib.svgdb_set_profiling(s._ctx, 1); lib.svgdb_reset_stats(s._ctx)
assert lib.svgdb_step(s._ctx, 10) == 0
st = s.Stats()
print("profiled 10 more steps: median %.3f grad %.3f phi %.3f (kernel %.3f) comm %.3f ms/step, passes %d" % (
    st["ms_median"] / 10, st["ms_grad"] / 10, st["ms_phi"] / 10, st.get("ms_phi_kernel", 0) / 10, st["ms_comm"] / 10, st["median_passes"]))
