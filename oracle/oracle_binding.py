"""ctypes binding of the CPU oracle (oracle/svgd_oracle.c).

TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` leg may import this module.  The product package svgdcpp_b200 never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libsvgd_oracle.so")

OPT_ADAGRAD, OPT_ADAM, OPT_RMSPROP = 0, 1, 2
SCALE_MEDIAN, SCALE_HESSIAN, SCALE_FIXED = 0, 1, 2

_dp = C.POINTER(C.c_double)


class _Config(C.Structure):
    _fields_ = [
        ("n", C.c_long), ("d", C.c_int), ("iters", C.c_int), ("n_components", C.c_int),
        ("means", _dp), ("covs", _dp), ("lse", C.c_int), ("scale_method", C.c_int),
        ("fixed_a", C.c_double), ("opt_kind", C.c_int),
        ("lr", C.c_double), ("beta1", C.c_double), ("beta2", C.c_double), ("eps", C.c_double),
        ("lb", _dp), ("ub", _dp),
    ]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "svgd_oracle.c")
    stale = (not os.path.exists(_LIB_PATH)) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src)
    if force or stale:
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.oracle_eigen_random.argtypes = [_dp, C.c_size_t, C.c_double, C.c_int, C.c_uint]
        L.oracle_eigen_random.restype = None
        L.oracle_lu_inverse.argtypes = [_dp, C.c_int, _dp]
        L.oracle_median.argtypes = [_dp, C.c_size_t]
        L.oracle_median.restype = C.c_double
        L.oracle_rbf_median_scale.argtypes = [_dp, C.c_long, C.c_int, _dp]
        L.oracle_rbf_median_scale.restype = C.c_double
        L.oracle_mvn_sum_logp_grad.argtypes = [_dp, C.c_long, C.c_int, C.c_int, _dp, _dp, C.c_int, _dp]
        L.oracle_phi.argtypes = [_dp, _dp, C.c_long, C.c_int, C.c_double, _dp]
        L.oracle_phi.restype = None
        L.oracle_rbf_hessian_scale.argtypes = [_dp, C.c_long, C.c_int, C.c_int, _dp, _dp, C.c_int, _dp]
        L.oracle_phi_matrix.argtypes = [_dp, _dp, C.c_long, C.c_int, _dp, _dp]
        L.oracle_phi_matrix.restype = None
        L.oracle_mvn_sum_logp.argtypes = [_dp, C.c_long, C.c_int, C.c_int, _dp, _dp, C.c_int, _dp]
        L.oracle_mvn_sum_logp.restype = C.c_int
        L.oracle_kernel_matrices.argtypes = [_dp, C.c_long, C.c_int, _dp, _dp, _dp]
        L.oracle_kernel_matrices.restype = None
        L.oracle_opt_step.argtypes = [C.c_int, C.c_size_t, _dp, C.c_double, C.c_double, C.c_double,
                                      C.c_double, C.POINTER(C.c_uint64), _dp, _dp, _dp]
        L.oracle_opt_step.restype = None
        L.oracle_clamp.argtypes = [_dp, C.c_long, C.c_int, _dp, _dp]
        L.oracle_clamp.restype = None
        L.oracle_svgd_run.argtypes = [C.POINTER(_Config), _dp, _dp, _dp]
        L.oracle_refshape_iterations_omp.argtypes = [C.POINTER(_Config), _dp, C.c_int, C.c_int]
        L.oracle_refshape_iterations_omp.restype = C.c_double
        L.oracle_blocked_iterations_omp.argtypes = [C.POINTER(_Config), _dp, C.c_int, C.c_int]
        L.oracle_blocked_iterations_omp.restype = C.c_double
        L.oracle_max_threads.restype = C.c_int
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(_dp) if a is not None else None


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def eigen_random(rows: int, cols: int, scale: float = 1.0, reseed: bool = True, seed: int = 1):
    """`scale * Eigen::MatrixXd::Random(rows, cols)`; returned particle-major (cols x rows)."""
    out = np.empty((cols, rows), dtype=np.float64)
    lib().oracle_eigen_random(_p(out), out.size, scale, int(reseed), seed)
    return out


def lu_inverse(A):
    A = _f64(A)
    out = np.empty_like(A)
    if lib().oracle_lu_inverse(_p(A), A.shape[0], _p(out)):
        raise ValueError("singular")
    return out


def median(v):
    v = _f64(v).copy().ravel()
    return lib().oracle_median(_p(v), v.size)


def rbf_median_scale(X):
    X = _f64(X)
    n, d = X.shape
    return lib().oracle_rbf_median_scale(_p(X), n, d, None)


def mvn_sum_logp_grad(X, means, covs, lse=False):
    X, means, covs = _f64(X), _f64(means), _f64(covs)
    n, d = X.shape
    means = means.reshape(-1, d)
    Cn = means.shape[0]
    covs = covs.reshape(Cn, d, d)
    G = np.empty_like(X)
    if lib().oracle_mvn_sum_logp_grad(_p(X), n, d, Cn, _p(means), _p(covs), int(lse), _p(G)):
        raise ValueError("singular covariance")
    return G


def mvn_sum_logp(X, means, covs, lse=False):
    """log p(x_i) of the sum of unnormalised Gaussians (Model::EvaluateLogModel)."""
    X = _f64(X)
    n, d = X.shape
    means, covs = _f64(np.atleast_2d(means)), _f64(np.asarray(covs).reshape(-1, d, d))
    out = np.empty(n)
    rc = lib().oracle_mvn_sum_logp(_p(X), n, d, means.shape[0], _p(means), _p(covs), int(lse), _p(out))
    if rc:
        raise RuntimeError("oracle_mvn_sum_logp failed (singular covariance?)")
    return out


def rbf_hessian_scale(X, means, covs, lse=False):
    """ScaleMethod::Hessian: A = 1/(2 d n) sum_i -Hessian(log p)(x_i), d x d."""
    X = _f64(X)
    n, d = X.shape
    means, covs = _f64(np.atleast_2d(means)), _f64(np.asarray(covs).reshape(-1, d, d))
    A = np.zeros((d, d))
    rc = lib().oracle_rbf_hessian_scale(_p(X), n, d, means.shape[0], _p(means), _p(covs), int(lse), _p(A))
    if rc:
        raise RuntimeError("oracle_rbf_hessian_scale failed")
    return A


def phi_matrix(X, G, A):
    X, G, A = _f64(X), _f64(G), _f64(A)
    n, d = X.shape
    out = np.empty_like(X)
    lib().oracle_phi_matrix(_p(X), _p(G), n, d, _p(A), _p(out))
    return out


def kernel_matrices(X, A):
    """kernel_matrix_ (n x n, [i, j] = k(x_j, x_i)) and kernel_grad_matrix_ ([i, j, :] = grad k(x_j, x_i)) of SVGD::ComputePhi."""
    X = _f64(X)
    n, d = X.shape
    A = _f64(np.asarray(A, dtype=np.float64) * np.eye(d) if np.ndim(A) == 0 else A)
    K = np.empty((n, n))
    dK = np.empty((n, n, d))
    lib().oracle_kernel_matrices(_p(X), n, d, _p(A), _p(K), _p(dK))
    return K, dK


def phi(X, G, a):
    X, G = _f64(X), _f64(G)
    n, d = X.shape
    out = np.empty_like(X)
    lib().oracle_phi(_p(X), _p(G), n, d, a, _p(out))
    return out


class OptState:
    def __init__(self, kind, shape, lr, beta1=0.9, beta2=0.999, eps=1e-8):
        self.kind, self.lr, self.beta1, self.beta2, self.eps = kind, lr, beta1, beta2, eps
        self.s1 = np.zeros(shape)
        self.s2 = np.zeros(shape)
        self.counter = C.c_uint64(0)

    def step(self, phi_mat):
        phi_mat = _f64(phi_mat)
        delta = np.empty_like(phi_mat)
        lib().oracle_opt_step(self.kind, phi_mat.size, _p(phi_mat), self.lr, self.beta1, self.beta2,
                              self.eps, C.byref(self.counter), _p(self.s1), _p(self.s2), _p(delta))
        return delta


def _make_cfg(n, d, iters, means, covs, lse, scale_method, fixed_a, opt_kind, lr, beta1, beta2, eps, lb, ub):
    means = _f64(means).reshape(-1, d)
    Cn = means.shape[0]
    covs = _f64(covs).reshape(Cn, d, d)
    lb = _f64(lb) if lb is not None else None
    ub = _f64(ub) if ub is not None else None
    cfg = _Config(n, d, iters, Cn, _p(means), _p(covs), int(lse), scale_method, fixed_a, opt_kind,
                  lr, beta1, beta2, eps, _p(lb), _p(ub))
    return cfg, (means, covs, lb, ub)  # keep arrays alive


def svgd_run(X0, iters, means, covs, *, opt_kind, lr, beta1=0.9, beta2=0.999, eps=1e-8,
             scale_method=SCALE_MEDIAN, fixed_a=0.0, lse=False, lb=None, ub=None,
             return_trace=False):
    """SVGD::Initialize + Run on a copy of X0 (n x d, particle-major).  Returns final X
    (and the per-iteration kernel scale `a` plus the last phi if return_trace)."""
    X = _f64(X0).copy()
    n, d = X.shape
    cfg, keep = _make_cfg(n, d, iters, means, covs, lse, scale_method, fixed_a, opt_kind, lr, beta1, beta2, eps, lb, ub)
    a_trace = np.zeros(max(iters, 1))
    phi_last = np.zeros_like(X)
    rc = lib().oracle_svgd_run(C.byref(cfg), _p(X), _p(a_trace), _p(phi_last))
    del keep
    if rc:
        raise RuntimeError("oracle_svgd_run failed")
    if return_trace:
        return X, a_trace[:iters], phi_last
    return X


def timed_iterations(X0, iters, means, covs, *, shape="refshape", threads=0, opt_kind=OPT_ADAM, lr=0.1,
                     beta1=0.9, beta2=0.999, eps=1e-8, scale_method=SCALE_MEDIAN, fixed_a=0.0, lse=True):
    """Seconds for `iters` CPU iterations on a copy of X0: 'refshape' (R) or 'blocked' (O)."""
    X = _f64(X0).copy()
    n, d = X.shape
    cfg, keep = _make_cfg(n, d, iters, means, covs, lse, scale_method, fixed_a, opt_kind, lr, beta1, beta2, eps, None, None)
    fn = lib().oracle_refshape_iterations_omp if shape == "refshape" else lib().oracle_blocked_iterations_omp
    secs = fn(C.byref(cfg), _p(X), iters, threads)
    del keep
    if secs < 0:
        raise MemoryError("oracle baseline allocation failed")
    return secs, X


def max_threads() -> int:
    return lib().oracle_max_threads()
