"""GPU parity tests of the tensor-core path for 64 < d <= 256 (kernels_phi_wide.cuh, kernels_dist_wide.cuh: one 128-particle
i-tile per CTA, k-chunked operands, Phi in column groups at d > 192) against the FP64 CPU oracle.  Tolerances as in
test_gpu_tc32.py: FAST 2e-4, PRECISE 1e-5 of max|phi|; kernel scale 1e-5; trajectories 1e-3 (RMS)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TC32 = 1
AUTO, FAST, PRECISE = 0, 1, 2
TOL = {FAST: 2e-4, PRECISE: 1e-5}


@pytest.fixture(scope="module")
def sv():
    import svgdcpp_b200

    svgdcpp_b200._capi.load()
    return svgdcpp_b200


def _problem(n, d, seed):
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((d, d))
    cov = A @ A.T / d + 0.5 * np.eye(d)
    mu = rng.standard_normal(d)
    x0 = np.asfortranarray(2.0 * rng.standard_normal((d, n)))
    return x0, mu, cov


@pytest.mark.parametrize("variant", [FAST, PRECISE])
@pytest.mark.parametrize("n,d", [(300, 128), (257, 100), (640, 192), (200, 130), (384, 256), (1000, 250), (129, 65)])
def test_wide_phi_fixed_scale(sv, oracle, n, d, variant):
    """The pair kernel alone (constant kernel scale: no distance pass)."""
    x0, mu, cov = _problem(n, d, seed=n + d)
    X = np.array(x0.T, order="C", copy=True)
    a_fix = 1.0 / (2.0 * d)
    model = sv.MultivariateNormal(mu, cov)
    svgd = sv.SVGD(d, 1, x0, sv.GaussianRBFKernel(x0, sv.ScaleMethod.Fixed, model, fixed_scale=a_fix), model, sv.AdaGrad(d, n, 0.1),
                   precision=TC32, tc32_variant=variant)
    phi, a = svgd.ComputePhi()
    svgd.close()
    G_ref = oracle.mvn_sum_logp_grad(X, mu[None], cov[None], lse=True)
    phi_ref = oracle.phi(X, G_ref, a_fix)
    err = np.max(np.abs(phi.T - phi_ref)) / np.max(np.abs(phi_ref))
    print("wide fixed-scale variant %d n=%d d=%d: phi max-rel err %.3g" % (variant, n, d, err))
    assert a == a_fix
    assert err < TOL[variant]


@pytest.mark.parametrize("n,d", [(300, 128), (777, 192), (513, 256), (1500, 200)])
def test_wide_median_scale_and_phi(sv, oracle, n, d):
    """The wide distance pass (exact median over all n^2 distances) and phi with it."""
    x0, mu, cov = _problem(n, d, seed=7 * n + d)
    X = np.array(x0.T, order="C", copy=True)
    model = sv.MultivariateNormal(mu, cov)
    svgd = sv.SVGD(d, 1, x0, sv.GaussianRBFKernel(x0, sv.ScaleMethod.Median, model), model, sv.AdaGrad(d, n, 0.1), precision=TC32,
                   tc32_variant=PRECISE)
    phi, a = svgd.ComputePhi()
    a_ref = oracle.rbf_median_scale(X)
    phi_ref = oracle.phi(X, oracle.mvn_sum_logp_grad(X, mu[None], cov[None], lse=True), a_ref)
    err = np.max(np.abs(phi.T - phi_ref)) / np.max(np.abs(phi_ref))
    print("wide median n=%d d=%d: a rel err %.3g, phi max-rel err %.3g" % (n, d, abs(a - a_ref) / a_ref, err))
    assert abs(a - a_ref) <= 1e-5 * a_ref
    assert err < 2e-5   # includes the scale's error (log(n) da on every kernel value)
    svgd.close()


@pytest.mark.parametrize("capacity", [None, 4096])
@pytest.mark.parametrize("opt", ["adagrad", "adam"])
def test_wide_trajectory(sv, oracle, opt, capacity, monkeypatch):
    """Steps with the median scale recomputed every iteration: bracket prediction (folded collecting passes), and with a small
    candidate buffer the histogram narrowing passes of the wide distance kernel."""
    if capacity is not None:
        monkeypatch.setenv("SVGDB_CAND_CAPACITY", str(capacity))
    n, d, iters = 700, 160, 12
    x0, mu, cov = _problem(n, d, seed=3)
    X0 = np.array(x0.T, order="C", copy=True)
    model = sv.MultivariateNormal(mu, cov)
    optimizer = sv.Adam(d, n, 0.1, 0.9, 0.999) if opt == "adam" else sv.AdaGrad(d, n, 0.1)
    svgd = sv.SVGD(d, iters, x0, sv.GaussianRBFKernel(x0, sv.ScaleMethod.Median, model), model, optimizer, precision=TC32)
    svgd.Initialize()
    svgd.Run()
    st = svgd.Stats()
    svgd.close()
    kind = oracle.OPT_ADAM if opt == "adam" else oracle.OPT_ADAGRAD
    ref = oracle.svgd_run(X0, iters, mu[None], cov[None], opt_kind=kind, lr=0.1, lse=True)
    rms = np.sqrt(np.mean((x0.T - ref) ** 2)) / np.sqrt(np.mean(ref ** 2))
    print("wide trajectory %s capacity %s: rms rel err %.3g, %d distance passes, %d bracket hits" % (opt, capacity, rms, st["median_passes"], st["median_bracket_hits"]))
    assert rms < 1e-3


def test_wide_config4_slice(sv, oracle):
    """BASELINE configs[3] on a slice the oracle can do (d = 256, 16 well-separated components, N = 2048) on the tensor-core path:
    kernel scale, mixture gradient, phi, then 3 AdaGrad steps."""
    from svgdcpp_b200 import synth

    n, d, C = 2048, 256, 16
    x0, means, covs = synth.gmm_problem(n, d, C)
    X0 = np.array(x0.T, order="C", copy=True)
    model = None
    for k in range(C):
        m = sv.MultivariateNormal(means[k], covs[k])
        model = m if model is None else model + m
    svgd = sv.SVGD(d, 3, x0, sv.GaussianRBFKernel(x0, sv.ScaleMethod.Median, model), model, sv.AdaGrad(d, n, 0.1), precision=TC32)
    phi, a = svgd.ComputePhi()
    a_ref = oracle.rbf_median_scale(X0)
    G_ref = oracle.mvn_sum_logp_grad(X0, means, covs, lse=True)
    phi_ref = oracle.phi(X0, G_ref, a_ref)
    e_phi = np.max(np.abs(phi.T - phi_ref)) / np.max(np.abs(phi_ref))
    svgd.Initialize()
    svgd.Run()
    svgd.close()
    ref = oracle.svgd_run(X0, 3, means, covs, opt_kind=oracle.OPT_ADAGRAD, lr=0.1, lse=True)
    rms = np.sqrt(np.mean((x0.T - ref) ** 2)) / np.sqrt(np.mean(ref ** 2))
    print("C4 slice on TC32 (n=%d d=%d C=%d): a rel err %.3g, phi max-rel err %.3g, 3-step rms rel err %.3g" % (n, d, C, abs(a - a_ref) / a_ref, e_phi, rms))
    assert abs(a - a_ref) <= 1e-5 * a_ref
    assert e_phi < 2e-5
    assert rms < 1e-4
