"""Sampled-row phi error at the headline shapes as a function of the TMEM accumulation length (SVGDB_PHI_MAX_SEG) and variant."""
import os, sys, subprocess, json
sys.path.insert(0, "."); sys.path.insert(0, "oracle")
import numpy as np
def run(workload, variant, max_seg):
    os.environ["SVGDB_PHI_MAX_SEG"] = str(max_seg)
    import ctypes as C
    import svgdcpp_b200 as sv
    from svgdcpp_b200 import synth, _capi
    if workload == "c3":
        n, d = 65536, 64
        x0, means, covs = synth.mvn_problem(n, d)
    elif workload.startswith("mvn"):  # mvn<d>: the config-3 recipe at another dimension
        n, d = 65536, workload[3:]  # mvn<d>[n<N>]
        if "n" in d:
            d, n = d.split("n"); n = int(n)
        d = int(d)
        x0, means, covs = synth.mvn_problem(n, d)
    else:
        n, d = 65536, 256
        x0, means, covs = synth.gmm_problem(n, d, 16)
    model = None
    for k in range(len(means)):
        m = sv.MultivariateNormal(means[k], covs[k]); model = m if model is None else model + m
    s = sv.SVGD(d, 1, x0, sv.GaussianRBFKernel(x0, sv.ScaleMethod.Median, model), model, sv.AdaGrad(d, n, 0.1), precision=1, tc32_variant=variant)
    G = s.EvaluateLogModelGrad().T
    phi, a = s.ComputePhi(); phi = phi.T
    X = np.array(x0.T, order="C")
    lib = _capi.load(); ms = C.c_float()
    lib.svgdb_time_kernel(s._ctx, 1, 3, 0, 0.0, C.byref(ms))
    s.close()
    scale = np.max(np.abs(phi)); worst = 0.0
    Xc = X - X.mean(0)
    rows = list(np.random.default_rng(0).integers(0, n, 6)) + [int(np.argmax(np.einsum("ij,ij->i", Xc, Xc)))]  # + the outermost particle
    for i in rows:
        diff = X - X[i]; k = np.exp(-a * np.einsum("ij,ij->i", diff, diff))
        ref = (k @ G + (-2 * a * diff * k[:, None]).sum(0)) / n
        worst = max(worst, np.max(np.abs(phi[i] - ref)) / scale)
    print("%s variant %d max_seg %4d: sampled phi rows max err / max|phi| = %.3g   pair pass %.3f ms" % (workload, variant, max_seg, worst, ms.value), flush=True)
if __name__ == "__main__":
    if len(sys.argv) > 1:
        run(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]))
    else:
        for wl, variant, segs in [("c3", 1, (100000, 128, 32)), ("c3", 2, (100000, 128, 32, 8)), ("c4s", 2, (100000, 256, 64, 16)), ("c4s", 1, (100000, 128))]:
            for ms in segs:
                subprocess.run([sys.executable, __file__, wl, str(variant), str(ms)])
