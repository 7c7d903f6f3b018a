"""C4 median trajectory: per-step kernel scale, distance passes and bracket hits (how well does the extrapolation predict the median?)."""
import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "oracle")
import numpy as np
import svgdcpp_b200 as sv
from svgdcpp_b200 import synth
n, d, C = int(sys.argv[1]) if len(sys.argv) > 1 else 65536, 256, 16
x0, means, covs = synth.gmm_problem(n, d, C)
model = None
for k in range(C):
    m = sv.MultivariateNormal(means[k], covs[k]); model = m if model is None else model + m
s = sv.SVGD(d, 1, x0, sv.GaussianRBFKernel(x0, sv.ScaleMethod.Median, model), model, sv.AdaGrad(d, n, 0.1), precision=1)
s.Initialize(); s._upload()
lib = s._lib
prev = None; passes = 0; hits = 0
for it in range(40):
    t0 = time.perf_counter()
    assert lib.svgdb_step(s._ctx, 1) == 0
    assert lib.svgdb_sync(s._ctx) == 0
    dt = time.perf_counter() - t0
    st = s.Stats()
    med2 = np.log(n) / st["last_scale"]
    print("step %2d: %.1f ms, passes %d, hit %d, median D2 %.8g, rel change %.3g" % (it, dt * 1e3, st["median_passes"] - passes, st["median_bracket_hits"] - hits, med2, 0 if prev is None else (med2 - prev) / prev), flush=True)
    prev = med2; passes = st["median_passes"]; hits = st["median_bracket_hits"]
s.close()
