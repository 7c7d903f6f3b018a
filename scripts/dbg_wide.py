import sys
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "oracle")
import svgdcpp_b200 as sv
import oracle_binding as oracle
for n, d in [(128, 256), (384, 256), (128, 250)]:
    rng = np.random.default_rng(n + d)
    A = rng.standard_normal((d, d)); cov = A @ A.T / d + 0.5 * np.eye(d); mu = rng.standard_normal(d)
    x0 = np.asfortranarray(2.0 * rng.standard_normal((d, n)))
    X = np.array(x0.T, order="C", copy=True)
    a_fix = 1.0 / (2.0 * d)
    model = sv.MultivariateNormal(mu, cov)
    for variant in (1, 2):
        s = sv.SVGD(d, 1, x0, sv.GaussianRBFKernel(x0, sv.ScaleMethod.Fixed, model, fixed_scale=a_fix), model, sv.AdaGrad(d, n, 0.1), precision=1, tc32_variant=variant)
        phi, a = s.ComputePhi(); s.close()
        phi = phi.T
        ref = oracle.phi(X, oracle.mvn_sum_logp_grad(X, mu[None], cov[None], lse=True), a_fix)
        bad = ~np.isfinite(phi)
        print("n=%d d=%d variant %d: non-finite %d of %d; by column block of 64: %s; by row block of 32: %s" % (
            n, d, variant, bad.sum(), bad.size, [int(bad[:, c:c + 64].sum()) for c in range(0, d, 64)], [int(bad[r:r + 32].sum()) for r in range(0, n, 32)]))
        ok = np.isfinite(phi)
        if ok.any():
            err = np.abs(phi - ref)
            print("   err/max|phi| on finite entries by column block: %s" % ["%.2e" % (np.max(np.where(ok[:, c:c + 64], err[:, c:c + 64], 0)) / np.max(np.abs(ref))) for c in range(0, d, 64)])
