"""SVGD::UpdateModelParameters / UpdateKernelParameters and MultivariateNormal::UpdateParameters between two Run() calls
(reference SVGD.hpp:304-332, Model/MultivariateNormal.hpp:94-115): the optimizer state and the particles stay on the device, the
target (or the constant kernel scale) changes.  Checked against the oracle's pieces composed step by step with the same change."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sv():
    import svgdcpp_b200

    svgdcpp_b200._capi.load()
    return svgdcpp_b200


def _oracle_steps(oracle, X, opt, iters, means, covs, fixed_a=None):
    """`iters` reference steps (SVGD.hpp:373-400) from the oracle's own pieces, with a caller-held optimizer state."""
    for _ in range(iters):
        a = oracle.rbf_median_scale(X) if fixed_a is None else fixed_a
        G = oracle.mvn_sum_logp_grad(X, means, covs, lse=True)
        X = X + opt.step(oracle.phi(X, G, a))
    return X


def _rel(a, b):
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))


@pytest.mark.parametrize("precision,tol", [(0, 1e-9), (1, 1e-3)])
def test_update_model_parameters_between_runs(sv, oracle, precision, tol):
    n, d, iters = 300, 6, 8
    rng = np.random.default_rng(5)
    mk = lambda: (lambda M: M @ M.T / d + 0.5 * np.eye(d))(rng.standard_normal((d, d)))
    mu1, cov1, mu2, cov2 = rng.standard_normal(d), mk(), rng.standard_normal(d) + 1.0, mk()
    x0 = np.asfortranarray(2.0 * rng.standard_normal((d, n)))
    X0 = np.array(x0.T, order="C", copy=True)
    model = sv.MultivariateNormal(mu1, cov1)
    svgd = sv.SVGD(d, iters, x0, sv.GaussianRBFKernel(x0, sv.ScaleMethod.Median, model), model, sv.Adam(d, n, 0.1, 0.9, 0.999),
                   precision=precision)
    svgd.Initialize()
    svgd.Run()
    svgd.UpdateModelParameters([mu2.reshape(-1, 1), cov2])          # SVGD.hpp:328-332 -> MultivariateNormal::UpdateParameters
    assert np.array_equal(model.GetParameters()[0].ravel(), mu2) and np.array_equal(model.GetParameters()[1], cov2)
    svgd.Run()
    opt = oracle.OptState(oracle.OPT_ADAM, X0.shape, 0.1)
    ref = _oracle_steps(oracle, X0, opt, iters, mu1[None], cov1[None])
    ref = _oracle_steps(oracle, ref, opt, iters, mu2[None], cov2[None])
    # the change must matter: without it the particles end somewhere else
    unchanged = _oracle_steps(oracle, _oracle_steps(oracle, X0, oracle.OptState(oracle.OPT_ADAM, X0.shape, 0.1), iters, mu1[None], cov1[None]),
                              oracle.OptState(oracle.OPT_ADAM, X0.shape, 0.1), iters, mu1[None], cov1[None])
    # FP64 mode: every coordinate (max norm); tensor-core mode: RMS, the norm its tolerance is stated in (Adam turns a coordinate
    # whose phi is below the arithmetic's noise into an O(lr) step of arbitrary sign, which a max norm then reports)
    err = _rel(x0.T, ref) if precision == 0 else float(np.sqrt(np.mean((x0.T - ref) ** 2)) / np.sqrt(np.mean(ref ** 2)))
    print("UpdateModelParameters, precision %d: final rel err %.3g (distance to the unchanged-model run %.3g)" % (precision, err, _rel(unchanged, ref)))
    assert _rel(unchanged, ref) > 1e-2
    assert err < tol
    # the model object itself: MultivariateNormal::UpdateParameters recomputes the normalisation constant (:182-186) and rejects bad shapes
    assert abs(model.GetNormalizationConstant() - 1.0 / ((2 * np.pi) ** (d / 2) * np.sqrt(np.linalg.det(cov2)))) < 1e-12 * model.GetNormalizationConstant()
    with pytest.raises(sv.DimensionMismatchException):
        model.UpdateParameters([np.zeros(d + 1), np.eye(d + 1)])
    with pytest.raises(sv.DimensionMismatchException):
        model.UpdateParameters([np.zeros(d), np.eye(d + 1)])
    svgd.close()


def test_update_kernel_parameters(sv, oracle):
    """A constant-scale kernel takes a new scale between runs; a median-scale kernel recomputes its scale at every step, so parameters
    pushed into it do not change the trajectory (GaussianRBFKernel::Step overwrites them, GaussianRBFKernel.hpp:141-156)."""
    n, d, iters = 120, 3, 6
    rng = np.random.default_rng(9)
    cov = (lambda M: M @ M.T / d + 0.5 * np.eye(d))(rng.standard_normal((d, d)))
    mu = rng.standard_normal(d)
    start = np.asfortranarray(1.5 * rng.standard_normal((d, n)))
    X0 = np.array(start.T, order="C", copy=True)
    # constant scale 0.8, then 0.3
    x0 = start.copy(order="F")
    model = sv.MultivariateNormal(mu, cov)
    svgd = sv.SVGD(d, iters, x0, sv.GaussianRBFKernel(x0, sv.ScaleMethod.Fixed, model, fixed_scale=0.8), model, sv.AdaGrad(d, n, 0.1))
    svgd.Initialize()
    svgd.Run()
    svgd.UpdateKernelParameters([0.3 * np.eye(d)])
    svgd.Run()
    with pytest.raises(ValueError):
        svgd.UpdateKernelParameters([np.diag(np.arange(1.0, d + 1.0))])   # not a * I: the device kernel has a scalar scale
    svgd.close()
    opt = oracle.OptState(oracle.OPT_ADAGRAD, X0.shape, 0.1)
    ref = _oracle_steps(oracle, _oracle_steps(oracle, X0, opt, iters, mu[None], cov[None], fixed_a=0.8), opt, iters, mu[None], cov[None], fixed_a=0.3)
    assert _rel(x0.T, ref) < 1e-9
    # median scale: the pushed parameters are overwritten by the next Step
    x1 = start.copy(order="F")
    svgd = sv.SVGD(d, iters, x1, sv.GaussianRBFKernel(x1, sv.ScaleMethod.Median, model), model, sv.AdaGrad(d, n, 0.1))
    svgd.Initialize()
    svgd.Run()
    svgd.UpdateKernelParameters([0.3 * np.eye(d)])
    svgd.Run()
    svgd.close()
    ref = oracle.svgd_run(X0, 2 * iters, mu[None], cov[None], opt_kind=oracle.OPT_ADAGRAD, lr=0.1)
    assert _rel(x1.T, ref) < 1e-9


def test_cpp_update_parameters(oracle, tmp_path):
    """The same through the C++ facade (tests/cpp/update_parameters.cpp)."""
    from test_facade_gpu import _build_and_run

    out = _build_and_run("update_parameters", tmp_path, src_dir=os.path.join("tests", "cpp"))
    rows = np.array([[float(t) for t in line.split()] for line in out.strip().splitlines()])
    d, n, iters = 2, 8, 5
    assert rows.shape == (2 * d, n)
    X0 = np.array([[1.5, -0.75, 0.25, 2.0, -1.25, 0.5, -2.0, 1.0], [-0.5, 1.0, 0.75, -1.5, 0.125, 2.25, 0.5, -1.0]]).T.copy()
    mu1, cov1 = np.array([[0.5, -0.25]]), np.array([[[0.5, 0.2], [0.2, 0.8]]])
    mu2, cov2 = np.array([[-1.0, 0.75]]), np.array([[[1.5, -0.3], [-0.3, 0.6]]])
    for scenario, (a1, a2) in enumerate([(None, None), (0.8, 0.3)]):
        opt = oracle.OptState(oracle.OPT_ADAM, X0.shape, 0.1)
        ref = _oracle_steps(oracle, X0, opt, iters, mu1, cov1, fixed_a=a1)
        ref = _oracle_steps(oracle, ref, opt, iters, mu2, cov2, fixed_a=a2)
        got = rows[scenario * d:(scenario + 1) * d].T
        print("C++ facade, scenario %d: final rel err %.3g" % (scenario, _rel(got, ref)))
        assert _rel(got, ref) < 1e-9
