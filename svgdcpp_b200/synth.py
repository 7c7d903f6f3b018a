"""Deterministic, portable synthetic inputs: splitmix64(seed, index) ->
53-bit uniform -> Box-Muller.  Pure numpy; used by bench.py and the tests, never by the kernels."""
from __future__ import annotations

import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x):
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = x
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return z ^ (z >> np.uint64(31))


def uniform(seed: int, count: int):
    """count doubles in (0, 1)."""
    with np.errstate(over="ignore"):
        idx = np.arange(count, dtype=np.uint64) + np.uint64(seed) * np.uint64(0x632BE59BD9B4E019)
    bits = _splitmix64(idx) >> np.uint64(11)
    return (bits.astype(np.float64) + 0.5) * (1.0 / 9007199254740992.0)


def normal(seed: int, shape):
    count = int(np.prod(shape))
    half = (count + 1) // 2
    u = uniform(seed, 2 * half)
    r = np.sqrt(-2.0 * np.log(u[:half]))
    t = 2.0 * np.pi * u[half:]
    z = np.concatenate([r * np.cos(t), r * np.sin(t)])[:count]
    return z.reshape(shape)


def uniform_pm1(seed: int, shape):
    return (2.0 * uniform(seed, int(np.prod(shape))) - 1.0).reshape(shape)


def mvn_problem(n: int, d: int):
    """Config 3 recipe: Sigma = A A^T / d + 0.5 I (A ~ N(0,1), seed 1), mu ~ N(0,1) (seed 2),
    X0 = 2 N(0,1) (seed 3).  Returns X0 as dim x n (reference layout), means (1 x d), covs (1 x d x d)."""
    A = normal(1, (d, d))
    cov = A @ A.T / d + 0.5 * np.eye(d)
    mu = normal(2, (d,))
    X0 = 2.0 * normal(3, (n, d))
    return np.asfortranarray(X0.T), mu[None, :], cov[None, :, :]


def gmm_problem(n: int, d: int, C: int):
    """Config 4 recipe: mu_k = 3 N(0,I) (seed 10+k), Sigma_k = A_k A_k^T / d + 0.5 I (seed 30+k),
    x_i = mu_{i mod C} + 1.5 chol(Sigma_{i mod C}) z_i (seed 50)."""
    means = np.stack([3.0 * normal(10 + k, (d,)) for k in range(C)])
    covs = []
    for k in range(C):
        A = normal(30 + k, (d, d))
        covs.append(A @ A.T / d + 0.5 * np.eye(d))
    covs = np.stack(covs)
    z = normal(50, (n, d))
    X0 = np.empty((n, d))
    for k in range(C):
        L = np.linalg.cholesky(covs[k])
        sel = np.arange(k, n, C)
        X0[sel] = means[k] + 1.5 * z[sel] @ L.T
    return np.asfortranarray(X0.T), means, covs
