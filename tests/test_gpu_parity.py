"""GPU parity tests (run with -m gpu on the B200 box): the CUDA path, called through the C ABI,
against the CPU oracle on identical seeded inputs, plus the reference's published example outputs.

Tolerances (FP64 mode): per-step phi and kernel scale <= 1e-11 relative; final particles after the
configured iteration count within 1e-9 of max|X|."""
import os

import numpy as np
import pytest

from helpers import assert_matches_printed, load_golden

pytestmark = pytest.mark.gpu

PHI_RTOL = 1e-11
FINAL_RTOL = 1e-9


@pytest.fixture(scope="module")
def sv():
    import svgdcpp_b200

    svgdcpp_b200._capi.load()  # fails loudly if the CUDA extension is missing
    return svgdcpp_b200


def _rel(a, b):
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


def _example(sv, oracle, g, precision=0):
    n, d = g["num_particles"], g["dim"]
    x0 = oracle.eigen_random(d, n, g["x0_scale"], reseed=True, seed=1).T.copy()  # dim x n
    model = None
    for mu, cov in zip(g["means"], g["covs"]):
        m = sv.MultivariateNormal(mu, cov)
        model = m if model is None else model + m
    kernel = sv.GaussianRBFKernel(x0, sv.ScaleMethod.Median, model)
    o = g["optimizer"]
    opt = sv.AdaGrad(d, n, o["lr"]) if o["kind"] == "adagrad" else sv.Adam(d, n, o["lr"], o["beta1"], o["beta2"])
    svgd = sv.SVGD(d, g["num_iterations"], x0, kernel, model, opt, precision=precision)
    svgd.Initialize()
    svgd.Run()
    svgd.close()
    return x0


@pytest.mark.parametrize("name", ["mvn_example", "gmm_example"])
def test_reference_examples_reproduced_on_gpu(sv, oracle, name):
    """mvn_example.cpp / gmm_example.cpp through the mirrored API: every printed digit of the
    reference's published final coordinates, and 1e-9 agreement with the oracle."""
    g = load_golden(name)
    xf = _example(sv, oracle, g)
    assert_matches_printed(xf.T, g["final"])
    x0 = oracle.eigen_random(g["dim"], g["num_particles"], g["x0_scale"], reseed=True, seed=1)
    o = g["optimizer"]
    kind = oracle.OPT_ADAGRAD if o["kind"] == "adagrad" else oracle.OPT_ADAM
    ref = oracle.svgd_run(x0, g["num_iterations"], g["means"], g["covs"], opt_kind=kind, lr=o["lr"],
                          beta1=o.get("beta1", 0.0), beta2=o.get("beta2", 0.0), eps=o["eps"])
    assert _rel(xf.T, ref) < FINAL_RTOL


def _mvn_setup(sv, n, d, seed=0, opt="adam", iters=1, **kw):
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((d, d))
    cov = A @ A.T / d + 0.5 * np.eye(d)
    mu = rng.standard_normal(d)
    x0 = np.asfortranarray(2.0 * rng.standard_normal((d, n)))
    model = sv.MultivariateNormal(mu, cov)
    kernel = sv.GaussianRBFKernel(x0, kw.pop("scale", sv.ScaleMethod.Median), model, fixed_scale=kw.pop("fixed_scale", 0.0))
    optimizer = {"adam": lambda: sv.Adam(d, n, 0.1, 0.9, 0.999), "adagrad": lambda: sv.AdaGrad(d, n, 0.1),
                 "rmsprop": lambda: sv.RMSProp(d, n, 0.05, 0.9)}[opt]()
    svgd = sv.SVGD(d, iters, x0, kernel, model, optimizer, **kw)
    return svgd, x0, mu[None], cov[None]


@pytest.mark.parametrize("n,d", [(1, 3), (2, 2), (10, 2), (63, 1), (64, 8), (65, 5), (200, 17), (300, 64), (257, 100), (130, 130), (96, 200)])
def test_phi_scale_and_grad_match_oracle(sv, oracle, n, d):
    """One ComputePhi (SVGD.hpp:407-454) on ragged / tiny / multi-chunk shapes."""
    svgd, x0, mu, cov = _mvn_setup(sv, n, d, seed=n * 1000 + d)
    X = np.array(x0.T, order="C", copy=True)
    G = svgd.EvaluateLogModelGrad().T
    G_ref = oracle.mvn_sum_logp_grad(X, mu, cov)
    assert _rel(G, G_ref) < 1e-12
    if n == 1:
        svgd.close()
        return  # median of a single zero distance: a = log(1)/0 -> NaN in the reference as well
    phi, a = svgd.ComputePhi()
    a_ref = oracle.rbf_median_scale(X)
    assert abs(a - a_ref) <= 1e-12 * a_ref
    phi_ref = oracle.phi(X, G_ref, a_ref)
    assert _rel(phi.T, phi_ref) < PHI_RTOL
    svgd.close()


@pytest.mark.parametrize("opt", ["adam", "adagrad", "rmsprop"])
def test_trajectory_matches_oracle(sv, oracle, opt):
    n, d, iters = 256, 64, 25
    svgd, x0, mu, cov = _mvn_setup(sv, n, d, seed=7, opt=opt, iters=iters)
    X0 = np.array(x0.T, order="C", copy=True)
    svgd.Initialize()
    svgd.Run()
    kind = {"adam": oracle.OPT_ADAM, "adagrad": oracle.OPT_ADAGRAD, "rmsprop": oracle.OPT_RMSPROP}[opt]
    lr = 0.05 if opt == "rmsprop" else 0.1
    ref = oracle.svgd_run(X0, iters, mu, cov, opt_kind=kind, lr=lr, beta1=0.9, beta2=0.999 if opt == "adam" else 0.0)
    assert _rel(x0.T, ref) < FINAL_RTOL
    st = svgd.Stats()
    assert st["iterations"] == iters and st["kernel_launches"] > 0
    svgd.close()


def test_run_is_resumable_and_in_place(sv, oracle):
    """Two Run() calls of k iterations equal one of 2k (state stays on the device between calls,
    x0 is updated in place like the reference's shared coordinate matrix, SVGD.hpp:393)."""
    n, d = 100, 6
    svgd, x0, mu, cov = _mvn_setup(sv, n, d, seed=11, iters=5)
    X0 = np.array(x0.T, order="C", copy=True)
    svgd.Initialize()
    svgd.Run()
    svgd.Run()
    ref = oracle.svgd_run(X0, 10, mu, cov, opt_kind=oracle.OPT_ADAM, lr=0.1)
    assert _rel(x0.T, ref) < FINAL_RTOL
    svgd.Initialize()  # zeroes the optimizer state again (Adam.hpp:61-67)
    X1 = np.array(x0.T, order="C", copy=True)
    svgd.Run()
    ref2 = oracle.svgd_run(X1, 5, mu, cov, opt_kind=oracle.OPT_ADAM, lr=0.1)
    assert _rel(x0.T, ref2) < FINAL_RTOL
    svgd.close()


def test_bounds_and_fixed_scale(sv, oracle):
    """Fixed-bandwidth kernel + box clamp (the tests/test_svgd.cpp configuration, with an MVN target)."""
    n, d, iters = 40, 2, 15
    svgd, x0, mu, cov = _mvn_setup(sv, n, d, seed=3, iters=iters, scale=sv.ScaleMethod.Fixed, fixed_scale=1.0,
                                   bound_lower=[-1.0, -0.5], bound_upper=[1.0, 0.75])
    X0 = np.array(x0.T, order="C", copy=True)
    svgd.Initialize()
    svgd.Run()
    ref = oracle.svgd_run(X0, iters, mu, cov, opt_kind=oracle.OPT_ADAM, lr=0.1, scale_method=oracle.SCALE_FIXED, fixed_a=1.0,
                          lb=[-1.0, -0.5], ub=[1.0, 0.75])
    assert np.max(x0[0]) <= 1.0 and np.min(x0[0]) >= -1.0 and np.max(x0[1]) <= 0.75 and np.min(x0[1]) >= -0.5
    assert np.sum(x0[0] == 1.0) + np.sum(x0[0] == -1.0) > 0  # the clamp is active
    assert _rel(x0.T, ref) < FINAL_RTOL
    svgd.close()


def test_device_hook_reproduces_reference_svgd_test(sv, oracle):
    """The reference's own SVGD test (tests/test_svgd.cpp:65-204): user model a cos(x0) + b cos(x1) + c x0 x1 + d through the
    device-gradient hook (the replacement for CppAD-taped Model lambdas), fixed-bandwidth kernel exp(-|x - x'|^2), Adam, box
    bounds, Eigen::MatrixXd::Random start, 15 iterations -- against the known answer of SURVEY.md section 8c item 3."""
    import ctypes as C

    import helpers

    hook_lib = C.CDLL(helpers.build_cos_hook())

    class CosParams(C.Structure):
        _fields_ = [("a", C.c_double), ("b", C.c_double), ("c", C.c_double), ("d", C.c_double)]

    params = CosParams(*helpers.COS_PARAMS)
    n, d, iters = 10, 2, 15
    x0 = np.asfortranarray(oracle.eigen_random(d, n, 1.0, reseed=True, seed=1).T)  # d x n, column-major like Eigen
    model = sv.Model(d)
    model.SetDeviceGradient(C.cast(hook_lib.cos_model_grad, C.c_void_p), C.cast(C.pointer(params), C.c_void_p))
    kernel = sv.GaussianRBFKernel(x0, sv.ScaleMethod.Fixed, model, fixed_scale=1.0)
    svgd = sv.SVGD(d, iters, x0, kernel, model, sv.Adam(d, n, 0.1, 0.9, 0.999), bound_lower=[-1.0, -1.0], bound_upper=[1.0, 1.0])
    svgd.Initialize()
    svgd.Run()
    svgd.close()
    err = max(np.max(np.abs(x0[0] - helpers.COS_KAT_ROW0)), np.max(np.abs(x0[1] - helpers.COS_KAT_ROW1)))
    print("device hook, test_svgd.cpp scenario: max abs err vs the known answer %.3g" % err)
    assert err < 5e-12  # the known answer carries 12 digits


def test_baseline_config_variants(sv, oracle):
    """BASELINE.json configs[0] / configs[1] as worded there (C1 = the MVN example's target with N=100 and Adam,
    C2 = three Gaussians, N=1000, AdaGrad), 1000 iterations, final particles against the oracle."""
    from svgdcpp_b200 import synth

    g1, g2 = load_golden("mvn_example"), load_golden("gmm_example")
    # C1: 2-D MVN target of mvn_example.cpp, 100 particles, X0 = 3 U(-1,1) (seed 1001), Adam(0.1, 0.9, 0.999)
    n, d, iters = 100, 2, 1000
    mu, cov = np.asarray(g1["means"][0], dtype=np.float64), np.asarray(g1["covs"][0], dtype=np.float64)
    x0 = np.asfortranarray(3.0 * synth.uniform_pm1(1001, (n, d)).T)
    X0 = np.array(x0.T, order="C", copy=True)
    model = sv.MultivariateNormal(mu, cov)
    svgd = sv.SVGD(d, iters, x0, sv.GaussianRBFKernel(x0, sv.ScaleMethod.Median, model), model, sv.Adam(d, n, 0.1, 0.9, 0.999))
    svgd.Initialize()
    svgd.Run()
    svgd.close()
    ref = oracle.svgd_run(X0, iters, mu[None], cov[None], opt_kind=oracle.OPT_ADAM, lr=0.1)
    print("C1 variant: final rel err %.3g" % _rel(x0.T, ref))
    assert _rel(x0.T, ref) < FINAL_RTOL
    # C2: the two components of gmm_example.cpp plus a third one, 1000 particles, X0 = 8 U(-1,1) (seed 1002), AdaGrad(0.1)
    n, iters = 1000, 1000
    means = np.array(list(g2["means"]) + [[-3.0, -3.5]], dtype=np.float64)
    covs = np.array(list(g2["covs"]) + [g1["covs"][0]], dtype=np.float64)
    x0 = np.asfortranarray(8.0 * synth.uniform_pm1(1002, (n, d)).T)
    X0 = np.array(x0.T, order="C", copy=True)
    model = None
    for k in range(3):
        m = sv.MultivariateNormal(means[k], covs[k])
        model = m if model is None else model + m
    svgd = sv.SVGD(d, iters, x0, sv.GaussianRBFKernel(x0, sv.ScaleMethod.Median, model), model, sv.AdaGrad(d, n, 0.1))
    svgd.Initialize()
    svgd.Run()
    svgd.close()
    ref = oracle.svgd_run(X0, iters, means, covs, opt_kind=oracle.OPT_ADAGRAD, lr=0.1, lse=True)
    print("C2 variant: final rel err %.3g" % _rel(x0.T, ref))
    assert _rel(x0.T, ref) < FINAL_RTOL


def test_config4_slice(sv, oracle):
    """BASELINE configs[3] on a slice the oracle can do (d=256, 16 components, N=1024): one ComputePhi, kernel scale and
    mixture gradient against the oracle (the far components underflow: log-sum-exp form), then 3 AdaGrad steps."""
    from svgdcpp_b200 import synth

    n, d, C = 1024, 256, 16
    x0, means, covs = synth.gmm_problem(n, d, C)
    X0 = np.array(x0.T, order="C", copy=True)
    model = None
    for k in range(C):
        m = sv.MultivariateNormal(means[k], covs[k])
        model = m if model is None else model + m
    svgd = sv.SVGD(d, 3, x0, sv.GaussianRBFKernel(x0, sv.ScaleMethod.Median, model), model, sv.AdaGrad(d, n, 0.1))
    phi, a = svgd.ComputePhi()
    a_ref = oracle.rbf_median_scale(X0)
    G_ref = oracle.mvn_sum_logp_grad(X0, means, covs, lse=True)
    phi_ref = oracle.phi(X0, G_ref, a_ref)
    print("C4 slice: a rel err %.3g, phi rel err %.3g" % (abs(a - a_ref) / a_ref, _rel(phi.T, phi_ref)))
    assert abs(a - a_ref) <= PHI_RTOL * a_ref
    assert _rel(phi.T, phi_ref) < 1e-10
    svgd.Initialize()
    svgd.Run()
    svgd.close()
    ref = oracle.svgd_run(X0, 3, means, covs, opt_kind=oracle.OPT_ADAGRAD, lr=0.1, lse=True)
    assert _rel(x0.T, ref) < FINAL_RTOL


def test_hessian_scale(sv, oracle):
    """ScaleMethod::Hessian (GaussianRBFKernel.hpp:189-210): scale matrix, phi and a short trajectory against
    the oracle, for one Gaussian (A = Sigma^-1 / (2 d)) and for a sum of two well-overlapping Gaussians.  The device path runs
    the scalar-bandwidth pair kernel on y = R x (A = R^T R), so the scale matrix must be positive definite."""
    rng = np.random.default_rng(5)
    for n, d, C in [(300, 8, 1), (257, 6, 2), (200, 64, 1)]:
        means = 0.4 * rng.standard_normal((C, d))
        covs = np.stack([(lambda M: M @ M.T / d + 0.7 * np.eye(d))(rng.standard_normal((d, d))) for _ in range(C)])
        x0 = np.asfortranarray(1.2 * rng.standard_normal((d, n)))
        X0 = np.array(x0.T, order="C", copy=True)
        model = None
        for k in range(C):
            m = sv.MultivariateNormal(means[k], covs[k])
            model = m if model is None else model + m
        svgd = sv.SVGD(d, 6, x0, sv.GaussianRBFKernel(x0, sv.ScaleMethod.Hessian, model), model, sv.Adam(d, n, 0.05, 0.9, 0.999))
        phi, _ = svgd.ComputePhi()
        A = svgd.GetScaleMatrix()
        A_ref = oracle.rbf_hessian_scale(X0, means, covs, lse=True)
        phi_ref = oracle.phi_matrix(X0, oracle.mvn_sum_logp_grad(X0, means, covs, lse=True), A_ref)
        print("hessian scale n=%d d=%d C=%d: A rel err %.3g, phi rel err %.3g" % (n, d, C, _rel(A, A_ref), _rel(phi.T, phi_ref)))
        assert _rel(A, A_ref) < 1e-12
        assert _rel(phi.T, phi_ref) < 1e-10
        svgd.Initialize()
        svgd.Run()
        svgd.close()
        ref = oracle.svgd_run(X0, 6, means, covs, opt_kind=oracle.OPT_ADAM, lr=0.05, scale_method=oracle.SCALE_HESSIAN, lse=True)
        assert _rel(x0.T, ref) < FINAL_RTOL


@pytest.mark.parametrize("method", ["median", "hessian", "fixed"])
def test_kernel_matrices_match_oracle(sv, oracle, method):
    """svgdb_compute_kernel_matrices (the matrices behind SVGDOptions::LogIntermediateMatrices, SVGD.hpp:434-448) against the oracle,
    and the reference's own assembly of phi from them (SVGD.hpp:453) against ComputePhi."""
    n, d = 37, 5
    rng = np.random.default_rng(17)
    cov = (lambda M: M @ M.T / d + 0.7 * np.eye(d))(rng.standard_normal((d, d)))
    mu = 0.3 * rng.standard_normal(d)
    x0 = np.asfortranarray(1.1 * rng.standard_normal((d, n)))
    X = np.array(x0.T, order="C", copy=True)
    model = sv.MultivariateNormal(mu, cov)
    scale = {"median": sv.ScaleMethod.Median, "hessian": sv.ScaleMethod.Hessian, "fixed": sv.ScaleMethod.Fixed}[method]
    kernel = sv.GaussianRBFKernel(x0, scale, model, fixed_scale=1.0)
    svgd = sv.SVGD(d, 1, x0, kernel, model, sv.AdaGrad(d, n, 0.1))
    if method == "fixed":
        svgd.UpdateKernelParameters([0.37 * np.eye(d)])
    K, dK = svgd.ComputeKernelMatrices()          # Eigen layouts: K[j, i], dK[j d + c, i]
    A = {"median": lambda: oracle.rbf_median_scale(X) * np.eye(d), "fixed": lambda: 0.37 * np.eye(d),
         "hessian": lambda: oracle.rbf_hessian_scale(X, mu[None], cov[None])}[method]()
    K_ref, dK_ref = oracle.kernel_matrices(X, A)
    assert _rel(K, K_ref.T) < 1e-12
    assert _rel(dK, dK_ref.reshape(n, n * d).T) < 1e-12
    G = svgd.EvaluateLogModelGrad()
    indexer = np.tile(np.eye(d), (1, n))          # kernel_grad_indexer_ (SVGD.hpp:250)
    phi_assembled = (G @ K + indexer @ dK) / n
    phi, _ = svgd.ComputePhi()
    assert _rel(phi_assembled, phi) < 1e-11
    svgd.close()


def test_kernel_matrices_size_limit(sv):
    """The inspection path refuses sizes whose matrices would not fit its 2 GiB budget."""
    n, d = 2200, 64  # 2200^2 x 65 doubles = 2.5 GB
    x0 = np.zeros((d, n), order="F")
    model = sv.MultivariateNormal(np.zeros(d), np.eye(d))
    svgd = sv.SVGD(d, 1, x0, sv.GaussianRBFKernel(x0, sv.ScaleMethod.Median, model), model, sv.AdaGrad(d, n, 0.1))
    with pytest.raises(sv.DimensionMismatchException, match="2 GiB"):
        svgd.ComputeKernelMatrices()
    svgd.close()


def test_python_log_intermediate_matrices(sv, oracle, tmp_path):
    """LogIntermediateMatrices through the Python mirror writes the reference's text layout."""
    n, d = 5, 2
    rng = np.random.default_rng(2)
    x0 = np.asfortranarray(rng.standard_normal((d, n)))
    model = sv.MultivariateNormal(np.zeros(d), np.eye(d))
    path = tmp_path / "log.txt"
    svgd = sv.SVGD(d, 2, x0, sv.GaussianRBFKernel(x0, sv.ScaleMethod.Median, model), model, sv.Adam(d, n, 0.1, 0.9, 0.999),
                   log_intermediate_matrices=True, intermediate_matrices_output_path=str(path))
    svgd.Initialize()
    svgd.Run()
    svgd.close()
    text = path.read_text()
    assert text.startswith("========== Step 1 ==========\nLogModelGrad=\n") and "========== Step 2 ==========" in text
    for name in ("\n\nKernel=\n", "\n\nKernelGrad=\n", "\n\nCoordMat=\n"):
        assert text.count(name) == 2
    last = text.rstrip("\n").split("CoordMat=\n")[-1]
    printed = np.array([[float(t) for t in line.split()] for line in last.splitlines()])
    assert np.allclose(printed, x0, rtol=1e-5, atol=1e-6)


def test_model_point_evaluations(sv, oracle):
    """Model::Evaluate{Model,LogModel,ModelGrad,LogModelGrad} and MultivariateNormal::*Normalized (Model.hpp:290-338,
    MultivariateNormal.hpp:143-175) run on the device; the values of all particles of a run come from svgdb_compute_log_model."""
    import ctypes as C

    d = 3
    rng = np.random.default_rng(8)
    mus = rng.standard_normal((2, d))
    covs = np.stack([(lambda M: M @ M.T / d + 0.6 * np.eye(d))(rng.standard_normal((d, d))) for _ in range(2)])
    mvn = sv.MultivariateNormal(mus[0], covs[0])
    mix = mvn + sv.MultivariateNormal(mus[1], covs[1])
    x = rng.standard_normal(d)
    for model, m, c in ((mvn, mus[:1], covs[:1]), (mix, mus, covs)):
        logp = oracle.mvn_sum_logp(x[None], m, c)[0]
        g = oracle.mvn_sum_logp_grad(x[None], m, c)[0]
        assert abs(model.EvaluateLogModel(x) - logp) <= 1e-13 * max(1.0, abs(logp))
        assert abs(model.EvaluateModel(x) - np.exp(logp)) <= 1e-13 * np.exp(logp)
        assert _rel(model.EvaluateLogModelGrad(x), g) < 1e-13
        assert _rel(model.EvaluateModelGrad(x), np.exp(logp) * g) < 1e-13
    # one Gaussian: the closed form, and the normalised density against scipy's definition
    diff = x - mus[0]
    q = diff @ np.linalg.solve(covs[0], diff)
    assert abs(mvn.EvaluateLogModel(x) + 0.5 * q) < 1e-13 * max(1.0, q)
    pdf = np.exp(-0.5 * q) / np.sqrt((2 * np.pi) ** d * np.linalg.det(covs[0]))
    assert abs(mvn.EvaluateModelNormalized(x) - pdf) <= 1e-13 * pdf
    assert abs(mvn.EvaluateLogModelNormalized(x) - np.log(pdf)) <= 1e-12
    assert _rel(mvn.EvaluateModelGradNormalized(x), -pdf * np.linalg.solve(covs[0], diff)) < 1e-12
    # every particle of a run in one call, far out where the literal log(sum exp) underflows: finite through log-sum-exp
    n = 50
    x0 = np.asfortranarray(rng.standard_normal((d, n)) + 60.0)
    svgd = sv.SVGD(d, 1, x0, sv.GaussianRBFKernel(x0, sv.ScaleMethod.Median, mix), mix, sv.AdaGrad(d, n, 0.1))
    svgd._upload()
    out = np.empty(n)
    assert svgd._lib.svgdb_compute_log_model(svgd._ctx, out.ctypes.data_as(C.POINTER(C.c_double))) == 0
    ref = oracle.mvn_sum_logp(np.ascontiguousarray(x0.T), mus, covs, lse=True)
    assert np.all(np.isfinite(out)) and _rel(out, ref) < 1e-13
    svgd.close()


def test_hessian_scale_rejects_indefinite_matrix(sv):
    """Far-apart components make the mean negative Hessian indefinite: the device path reports it instead of running a kernel that
    is not positive definite (the reference would go on with exp(-d^T A d) > 1)."""
    d, n = 2, 64
    rng = np.random.default_rng(9)
    x0 = np.asfortranarray(0.3 * rng.standard_normal((d, n)))
    model = sv.MultivariateNormal([4.0, 0.0], 0.5 * np.eye(d)) + sv.MultivariateNormal([-4.0, 0.0], 0.5 * np.eye(d))
    svgd = sv.SVGD(d, 1, x0, sv.GaussianRBFKernel(x0, sv.ScaleMethod.Hessian, model), model, sv.AdaGrad(d, n, 0.1))
    with pytest.raises(RuntimeError, match="not positive definite"):
        svgd.ComputePhi()
    svgd.close()


def test_mixture_gradient_log_sum_exp(sv, oracle):
    """16-D, 5-component sum of Gaussians incl. far-away components (config-4 style)."""
    from svgdcpp_b200 import synth

    n, d, C = 333, 16, 5
    x0, means, covs = synth.gmm_problem(n, d, C)
    model = None
    for k in range(C):
        m = sv.MultivariateNormal(means[k], covs[k])
        model = m if model is None else model + m
    kernel = sv.GaussianRBFKernel(x0, sv.ScaleMethod.Median, model)
    svgd = sv.SVGD(d, 10, x0, kernel, model, sv.AdaGrad(d, n, 0.1))
    X0 = np.array(x0.T, order="C", copy=True)
    G = svgd.EvaluateLogModelGrad().T
    G_ref = oracle.mvn_sum_logp_grad(X0, means, covs, lse=True)
    assert _rel(G, G_ref) < 1e-12
    svgd.Initialize()
    svgd.Run()
    ref = oracle.svgd_run(X0, 10, means, covs, opt_kind=oracle.OPT_ADAGRAD, lr=0.1, lse=True)
    assert _rel(x0.T, ref) < FINAL_RTOL
    svgd.close()


@pytest.mark.parametrize("n,d,C", [(333, 16, 5), (1000, 64, 3), (700, 5, 2), (2048, 256, 16), (513, 100, 1)])
def test_mixture_gradient_through_library_gemm(sv, oracle, monkeypatch, n, d, C):
    """The mixture gradient as C DGEMMs Y = (X - mu_c) Sigma_c^-1 + streaming log-sum-exp kernels (launch_grad_gemm; automatic at
    d >= 32 and n C >= 16,384, forced here): same values as the oracle and as the one-kernel form, incl. far components, a
    dimension that is no multiple of anything, one component, and 10 steps of the whole path on top."""
    from svgdcpp_b200 import synth

    x0, means, covs = synth.gmm_problem(n, d, C)
    X0 = np.array(x0.T, order="C", copy=True)
    G_ref = oracle.mvn_sum_logp_grad(X0, means, covs, lse=True)
    got = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("SVGDB_GRAD_GEMM", mode)
        model = None
        for k in range(C):
            m = sv.MultivariateNormal(means[k], covs[k])
            model = m if model is None else model + m
        x = x0.copy(order="F")
        svgd = sv.SVGD(d, 10, x, sv.GaussianRBFKernel(x, sv.ScaleMethod.Median, model), model, sv.AdaGrad(d, n, 0.1))
        got[mode] = svgd.EvaluateLogModelGrad().T.copy()
        if mode == "1" and n * d <= 64 * 1024:
            svgd.Initialize()
            svgd.Run()
            ref = oracle.svgd_run(X0, 10, means, covs, opt_kind=oracle.OPT_ADAGRAD, lr=0.1, lse=True)
            assert _rel(x.T, ref) < FINAL_RTOL
        svgd.close()
    print("n=%d d=%d C=%d: gemm form vs oracle %.3g, one-kernel form vs oracle %.3g, between them %.3g"
          % (n, d, C, _rel(got["1"], G_ref), _rel(got["0"], G_ref), _rel(got["1"], got["0"])))
    assert _rel(got["1"], G_ref) < 1e-12
    assert _rel(got["1"], got["0"]) < 1e-12


def test_mixture_finite_where_reference_underflows(sv):
    """Every exp(-q/2) underflows in double: the reference's log(sum exp) is NaN;
    the device path evaluates the same gradient through log-sum-exp and stays finite."""
    d = 4
    x0 = np.asfortranarray(np.full((d, 8), 60.0) + np.arange(8)[None, :])
    for mode in ("0", "1"):  # the one-kernel form and the library-GEMM form
        os.environ["SVGDB_GRAD_GEMM"] = mode
        try:
            model = sv.MultivariateNormal(np.zeros(d), np.eye(d)) + sv.MultivariateNormal(np.ones(d), np.eye(d))
            kernel = sv.GaussianRBFKernel(x0, sv.ScaleMethod.Median, model)
            svgd = sv.SVGD(d, 1, x0, kernel, model, sv.AdaGrad(d, 8, 0.1))
            G = svgd.EvaluateLogModelGrad()
            assert np.all(np.isfinite(G))
            # dominated by the nearer component (mean 1): grad ~ -(x - 1)
            assert np.allclose(G, -(x0 - 1.0), rtol=1e-9)
            svgd.close()
        finally:
            del os.environ["SVGDB_GRAD_GEMM"]


@pytest.mark.parametrize("capacity", [64, 1000])
def test_median_select_narrowing_paths(sv, oracle, capacity, monkeypatch):
    """Tiny candidate buffers force the histogram-narrowing passes, the bracket prediction and the
    tie handling of the exact on-device select (GaussianRBFKernel.hpp:222-254 semantics)."""
    monkeypatch.setenv("SVGDB_CAND_CAPACITY", str(capacity))
    rng = np.random.default_rng(5)
    for n, d in [(90, 3), (151, 64), (64, 2)]:
        svgd, x0, mu, cov = _mvn_setup(sv, n, d, seed=n)
        X = np.array(x0.T, order="C", copy=True)
        a = svgd.ComputeScale()
        a_ref = oracle.rbf_median_scale(X)
        assert abs(a - a_ref) <= 1e-12 * a_ref, (n, d)
        # several steps: the predicted-bracket path must keep agreeing with the oracle
        svgd.Initialize()
        svgd.Step(6)
        ref = oracle.svgd_run(X, 6, mu, cov, opt_kind=oracle.OPT_ADAM, lr=0.1)
        assert _rel(x0.T, ref) < FINAL_RTOL
        assert svgd.Stats()["median_passes"] >= 6
        svgd.close()
    # heavy ties: many duplicated particles (more equal distances than the buffer holds)
    n, d = 80, 3
    base = rng.standard_normal((4, d))
    X = np.ascontiguousarray(base[rng.integers(0, 4, n)])
    x0 = np.asfortranarray(X.T)
    model = sv.MultivariateNormal(np.zeros(d), np.eye(d))
    svgd = sv.SVGD(d, 1, x0, sv.GaussianRBFKernel(x0, sv.ScaleMethod.Median, model), model, sv.AdaGrad(d, n, 0.1))
    D = np.sqrt(((X[:, None] - X[None]) ** 2).sum(-1))
    med = np.median(D.ravel())
    a = svgd.ComputeScale()
    if med > 0:
        assert abs(a - np.log(n) / med ** 2) <= 1e-10 * a
    else:
        assert np.isinf(a)
    svgd.close()


def test_error_mapping(sv):
    x0 = np.zeros((3, 5), order="F")
    model = sv.MultivariateNormal(np.zeros(3), np.eye(3))
    kernel = sv.GaussianRBFKernel(x0, sv.ScaleMethod.Median, model)
    with pytest.raises(sv.DimensionMismatchException):
        sv.SVGD(2, 1, x0, kernel, model, sv.AdaGrad(3, 5, 0.1))
    with pytest.raises(sv.DimensionMismatchException):
        sv.SVGD(3, 1, x0, kernel, model, sv.AdaGrad(3, 5, 0.1), bound_lower=[0.0, 0.0], bound_upper=[1.0, 1.0])
    with pytest.raises(ValueError):
        sv.SVGD(3, 1, x0, None, model, sv.AdaGrad(3, 5, 0.1))
    with pytest.raises(ValueError):
        sv.Adam(3, 5, 0.1, 1.0, 0.999)
    with pytest.raises(sv.UnsetException):
        sv.SVGD(3, 1, x0, kernel, sv.Model(3), sv.AdaGrad(3, 5, 0.1))
    with pytest.raises(RuntimeError):  # singular covariance
        sv.SVGD(3, 1, x0, kernel, sv.MultivariateNormal(np.zeros(3), np.zeros((3, 3))), sv.AdaGrad(3, 5, 0.1))


def test_full_size_properties(sv):
    """BASELINE config 3 (N=65,536, d=64): size-independent checks where the oracle cannot run.
    (1) sampled rows of phi recomputed in O(N d) on the host from the device's own scale;
    (2) the scale is log(N)/med^2 with about half of sampled distances below med."""
    from svgdcpp_b200 import synth

    n, d = 65536, 64
    x0, means, covs = synth.mvn_problem(n, d)
    model = sv.MultivariateNormal(means[0], covs[0])
    kernel = sv.GaussianRBFKernel(x0, sv.ScaleMethod.Median, model)
    svgd = sv.SVGD(d, 1, x0, kernel, model, sv.Adam(d, n, 0.1, 0.9, 0.999))
    X = np.array(x0.T, order="C", copy=True)
    phi, a = svgd.ComputePhi()
    G = -(X - means[0]) @ np.linalg.inv(covs[0])
    rows = np.random.default_rng(0).integers(0, n, 12)
    for i in rows:
        d2 = ((X - X[i]) ** 2).sum(1)
        k = np.exp(-a * d2)
        ref = (k @ G + (-2 * a * (X - X[i]) * k[:, None]).sum(0)) / n
        assert np.max(np.abs(phi[:, i] - ref)) <= 1e-10 * np.max(np.abs(ref))
    # the scale is log(n)/med^2 with med the median distance: half of 4M random ordered pairs lie below it
    med = np.sqrt(np.log(n) / a)
    rng = np.random.default_rng(1)
    ii, jj = rng.integers(0, n, 4_000_000), rng.integers(0, n, 4_000_000)
    below = 0
    for s0 in range(0, ii.size, 500_000):
        sl = slice(s0, s0 + 500_000)
        below += int(np.sum(np.sqrt(((X[ii[sl]] - X[jj[sl]]) ** 2).sum(1)) < med))
    assert abs(below / ii.size - 0.5) < 2e-3   # 8 sigma of the sampling error
    # one full step moves every particle by at most lr (Adam's first step is lr * sign(phi))
    before = x0.copy()
    svgd.Initialize()
    svgd.Step(1)
    step = np.abs(x0 - before)
    assert np.all(np.isfinite(x0)) and np.max(step) <= 0.1 * (1 + 1e-6) and np.mean(step) > 0.05
    svgd.close()
