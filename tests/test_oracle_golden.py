"""Pins the CPU oracle against the reference's own published outputs.

The reference prints Eigen matrices with 6 significant digits; every printed digit of the
initial AND final particle sets of both shipped examples must be reproduced.
"""
import numpy as np
import pytest

from helpers import assert_matches_printed, load_golden


def _run_example(oracle, g):
    n, d = g["num_particles"], g["dim"]
    x0 = oracle.eigen_random(d, n, g["x0_scale"], reseed=True, seed=1)  # unseeded rand() == srand(1)
    opt = g["optimizer"]
    kind = {"adagrad": oracle.OPT_ADAGRAD, "adam": oracle.OPT_ADAM}[opt["kind"]]
    xf = oracle.svgd_run(x0, g["num_iterations"], g["means"], g["covs"], opt_kind=kind, lr=opt["lr"],
                         beta1=opt.get("beta1", 0.0), beta2=opt.get("beta2", 0.0), eps=opt["eps"])
    return x0, xf


@pytest.mark.parametrize("name", ["mvn_example", "gmm_example"])
def test_oracle_reproduces_reference_example(oracle, name):
    g = load_golden(name)
    x0, xf = _run_example(oracle, g)
    assert_matches_printed(x0, g["initial"])
    assert_matches_printed(xf, g["final"])


def test_eigen_random_first_value(oracle):
    # the well-known first value of an unseeded Eigen::MatrixXd::Random
    x = oracle.eigen_random(2, 1, 1.0)
    assert abs(x[0, 0] - 0.680375434309419) < 1e-15


def test_median_conventions(oracle):
    # even count: mean of the two middle order statistics; odd: the middle one
    assert oracle.median([4.0, 1.0, 3.0, 2.0]) == 2.5
    assert oracle.median([5.0, 1.0, 3.0]) == 3.0
    rng = np.random.default_rng(0)
    for n in (2, 7, 100, 1001, 4096):
        v = rng.standard_normal(n)
        assert oracle.median(v) == np.median(v)


def test_median_scale_definition(oracle):
    # a = log(n) / median(all n*n distances, zeros and both orderings included)^2
    rng = np.random.default_rng(1)
    X = rng.standard_normal((37, 5))
    D = np.sqrt(np.maximum(((X[:, None, :] - X[None, :, :]) ** 2).sum(-1), 0.0))
    a_np = np.log(37) / np.median(D.ravel()) ** 2
    assert abs(oracle.rbf_median_scale(X) - a_np) < 1e-12 * a_np


def test_phi_matches_gram_form(oracle):
    # eq. 8 literal == (1/n)[K^T (G - 2aX) + 2a X rowsum(K)]  (the form the CUDA kernels use)
    rng = np.random.default_rng(2)
    n, d, a = 50, 7, 0.37
    X = rng.standard_normal((n, d))
    G = rng.standard_normal((n, d))
    D2 = ((X[:, None, :] - X[None, :, :]) ** 2).sum(-1)
    K = np.exp(-a * D2)
    ref = (K @ (G - 2 * a * X) + 2 * a * X * K.sum(1, keepdims=True)) / n
    got = oracle.phi(X, G, a)
    assert np.max(np.abs(got - ref)) < 1e-14 * np.max(np.abs(ref)) * 10


def test_mvn_grad_closed_form(oracle):
    rng = np.random.default_rng(3)
    d = 6
    A = rng.standard_normal((d, d))
    cov = A @ A.T / d + 0.5 * np.eye(d)
    mu = rng.standard_normal(d)
    X = rng.standard_normal((20, d))
    G = oracle.mvn_sum_logp_grad(X, mu[None], cov[None])
    ref = -(X - mu) @ np.linalg.inv(cov)
    assert np.max(np.abs(G - ref)) < 1e-12
    # mixture: direct log(sum exp) and log-sum-exp forms agree where both are finite
    mus = rng.standard_normal((3, d))
    covs = np.stack([cov, 2 * cov, cov + np.eye(d)])
    G1 = oracle.mvn_sum_logp_grad(X, mus, covs, lse=False)
    G2 = oracle.mvn_sum_logp_grad(X, mus, covs, lse=True)
    assert np.max(np.abs(G1 - G2)) < 1e-13
    h = np.stack([-0.5 * np.einsum("ni,ij,nj->n", X - m, np.linalg.inv(c), X - m) for m, c in zip(mus, covs)], 1)
    w = np.exp(h - h.max(1, keepdims=True))
    w /= w.sum(1, keepdims=True)
    ref = sum(w[:, [k]] * (-(X - mus[k]) @ np.linalg.inv(covs[k])) for k in range(3))
    assert np.max(np.abs(G2 - ref)) < 1e-12


def test_optimizers(oracle):
    rng = np.random.default_rng(4)
    phi = rng.standard_normal((5, 3))
    st = oracle.OptState(oracle.OPT_ADAM, phi.shape, 0.1, 0.9, 0.999, 1e-8)
    m = v = 0
    for t in range(1, 4):
        delta = st.step(phi * t)
        m = 0.9 * m + 0.1 * phi * t
        v = 0.999 * v + 0.001 * (phi * t) ** 2
        ref = 0.1 * (m / (1 - 0.9 ** t)) / (1e-8 + np.sqrt(v / (1 - 0.999 ** t)))  # eps OUTSIDE the sqrt
        assert np.allclose(delta, ref, rtol=1e-14, atol=0)
    st = oracle.OptState(oracle.OPT_ADAGRAD, phi.shape, 0.1)
    s = 0
    for t in range(1, 4):
        delta = st.step(phi)
        s = s + phi ** 2
        assert np.allclose(delta, 0.1 * phi / (1e-8 + np.sqrt(s)), rtol=1e-14, atol=0)
    st = oracle.OptState(oracle.OPT_RMSPROP, phi.shape, 0.1, beta1=0.9)
    s = 0
    for t in range(1, 4):
        delta = st.step(phi)
        s = 0.9 * s + 0.1 * phi ** 2
        assert np.allclose(delta, 0.1 * phi / (1e-8 + np.sqrt(s)), rtol=1e-14, atol=0)


def test_cpu_baselines_agree_with_literal(oracle):
    # the two timed CPU variants are the same algorithm as the literal oracle
    rng = np.random.default_rng(5)
    n, d = 64, 4
    X0 = 2 * rng.standard_normal((n, d))
    mu = rng.standard_normal((1, d))
    cov = np.eye(d)[None] * 1.5
    ref = oracle.svgd_run(X0, 3, mu, cov, opt_kind=oracle.OPT_ADAM, lr=0.1, lse=True)
    for shape in ("refshape", "blocked"):
        secs, X = oracle.timed_iterations(X0, 3, mu, cov, shape=shape, threads=2)
        assert secs >= 0
        assert np.max(np.abs(X - ref)) < 1e-10


def test_reference_svgd_test_scenario_known_answer(oracle):
    """tests/test_svgd.cpp:65-204 of the reference (cosine user model, fixed-bandwidth kernel, Adam, box bounds,
    Eigen::MatrixXd::Random start, 15 iterations): the oracle's phi / optimizer restatement reproduces the known answer
    of SURVEY.md section 8c item 3 to its 12 printed digits."""
    import helpers

    X = np.array(oracle.eigen_random(2, 10, 1.0, reseed=True, seed=1), order="C", copy=True)  # particle-major 10 x 2
    opt = oracle.OptState(oracle.OPT_ADAM, X.shape, 0.1)
    for _ in range(15):
        X = X + opt.step(oracle.phi(X, helpers.cos_model_grad(X), 1.0))
        X = np.maximum(np.minimum(X, 1.0), -1.0)  # min-then-max clamp of SVGD.hpp:396-399
    assert np.max(np.abs(X[:, 0] - helpers.COS_KAT_ROW0)) < 5e-12
    assert np.max(np.abs(X[:, 1] - helpers.COS_KAT_ROW1)) < 5e-12


def test_hessian_scale_closed_form_matches_finite_differences(oracle):
    """ScaleMethod::Hessian (GaussianRBFKernel.hpp:189-210) is not pinned by any reference output; the oracle's closed-form
    Hessian of log p for sums of Gaussians is checked against central differences of its own gradient, the single-Gaussian
    case against A = Sigma^-1 / (2 d), and the matrix-scale phi against the scalar one for A = a I."""
    rng = np.random.default_rng(11)
    d, n, C = 5, 40, 3
    means = rng.standard_normal((C, d)) * 1.5
    covs = np.stack([(lambda M: M @ M.T / d + 0.5 * np.eye(d))(rng.standard_normal((d, d))) for _ in range(C)])
    X = rng.standard_normal((n, d)) * 1.5
    A = oracle.rbf_hessian_scale(X, means, covs, lse=True)
    H, eps = np.zeros((d, d)), 1e-5
    for k in range(d):
        Xp, Xm = X.copy(), X.copy()
        Xp[:, k] += eps
        Xm[:, k] -= eps
        H[:, k] = -((oracle.mvn_sum_logp_grad(Xp, means, covs, lse=True) - oracle.mvn_sum_logp_grad(Xm, means, covs, lse=True)) / (2 * eps)).sum(0)
    assert np.max(np.abs(A - H / (2 * d * n))) < 1e-8 * np.max(np.abs(A))
    assert np.max(np.abs(oracle.rbf_hessian_scale(X, means[:1], covs[:1]) - np.linalg.inv(covs[0]) / (2 * d))) < 1e-14
    G = oracle.mvn_sum_logp_grad(X, means, covs, lse=True)
    assert np.max(np.abs(oracle.phi_matrix(X, G, 0.37 * np.eye(d)) - oracle.phi(X, G, 0.37))) < 1e-15


def test_oracle_kernel_matrices_assemble_phi(oracle):
    """The oracle's kernel / kernel-gradient matrices (SVGD.hpp:434-448), pushed through the reference's own assembly
    phi = (G K + indexer dK) / n (SVGD.hpp:453), give the oracle's phi: scalar and matrix scale."""
    n, d = 23, 4
    rng = np.random.default_rng(4)
    X = rng.standard_normal((n, d))
    G = rng.standard_normal((n, d))
    M = rng.standard_normal((d, d))
    for A, phi_ref in ((0.6 * np.eye(d), oracle.phi(X, G, 0.6)), (M @ M.T / d + 0.3 * np.eye(d), None)):
        K, dK = oracle.kernel_matrices(X, A)      # K[i, j] = k(x_j, x_i); dK[i, j, :] = grad k(x_j, x_i)
        phi = (np.einsum("ij,jc->ic", K, G) + dK.sum(axis=1)) / n
        if phi_ref is None:
            phi_ref = oracle.phi_matrix(X, G, A)
        assert np.max(np.abs(phi - phi_ref)) <= 1e-13 * np.max(np.abs(phi_ref))
        assert np.allclose(np.diag(K), 1.0) and np.allclose(K, K.T)


def test_oracle_logp_is_consistent_with_its_gradient(oracle):
    """Central differences of the oracle's log p reproduce the oracle's grad log p (sum of two Gaussians)."""
    d = 3
    rng = np.random.default_rng(12)
    mus = rng.standard_normal((2, d))
    covs = np.stack([(lambda M: M @ M.T / d + 0.6 * np.eye(d))(rng.standard_normal((d, d))) for _ in range(2)])
    x = rng.standard_normal(d)
    g = oracle.mvn_sum_logp_grad(x[None], mus, covs)[0]
    h = 1e-6
    fd = np.array([(oracle.mvn_sum_logp((x + h * e)[None], mus, covs)[0] - oracle.mvn_sum_logp((x - h * e)[None], mus, covs)[0]) / (2 * h)
                   for e in np.eye(d)])
    assert np.max(np.abs(fd - g)) < 1e-8 * max(1.0, np.max(np.abs(g)))
