"""Shared test helpers: golden loading, 6-significant-digit printing like Eigen's operator<<,
and the deterministic synthetic inputs of SURVEY.md section 8(d)."""
import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    with open(os.path.join(GOLDEN_DIR, name + ".json")) as f:
        return json.load(f)


def sig6(x):
    """Eigen's default stream precision (6 significant digits, %g-like)."""
    return float("%.6g" % x)


def assert_matches_printed(values, printed, digits=6):
    values = np.asarray(values, dtype=np.float64).ravel()
    printed = np.asarray(printed, dtype=np.float64).ravel()
    assert values.shape == printed.shape
    fmt = "%%.%dg" % digits
    for v, p in zip(values, printed):
        assert float(fmt % v) == p, "value %r prints as %s, reference printed %r" % (v, fmt % v, p)


# ---- the reference's own SVGD test scenario (tests/test_svgd.cpp:65-204), restated as a known-answer test ----------
# p(x) = 7.5 cos(x0) + 10 cos(x1) + 3 x0 x1 - 6 (user model), k = exp(-|x - x'|^2) (fixed bandwidth a = 1),
# Adam(0.1, 0.9, 0.999), bounds [-1, 1]^2, X0 = Eigen::MatrixXd::Random(2, 10) (unseeded glibc rand()), 15 iterations.
# Final particles from SURVEY.md section 8c item 3 (numpy restatement of the reference's manual computation).
COS_PARAMS = (7.5, 10.0, 3.0, -6.0)
COS_KAT_ROW0 = [1, 1, 0.311194941729, -0.470273792107, -1, -0.244114628268, -0.453814082155, 1, -0.016266095191, -1]
COS_KAT_ROW1 = [-0.270291345314, 0.555114197298, -1, 0.836541188037, -0.791217316816, 0.039423964202, -1, 1, 0.440006548476,
                -0.308552375847]


def cos_model_grad(X):
    """grad log p of the cosine model, X is n x 2 (reference: log_model_grad_fun, tests/test_svgd.cpp:157-170)."""
    a, b, c, d = COS_PARAMS
    den = a * np.cos(X[:, 0]) + b * np.cos(X[:, 1]) + c * X[:, 0] * X[:, 1] + d
    return np.stack([(-a * np.sin(X[:, 0]) + c * X[:, 1]) / den, (-b * np.sin(X[:, 1]) + c * X[:, 0]) / den], axis=1)


def build_cos_hook():
    """nvcc-builds tests/cuda/cos_model_hook.cu (the device-gradient hook of the cosine model) in-tree; returns the .so path."""
    import subprocess

    here = os.path.dirname(os.path.abspath(__file__))
    src = os.path.join(here, "cuda", "cos_model_hook.cu")
    out_dir = os.path.join(here, "cuda", "_build")
    out = os.path.join(out_dir, "libcos_hook.so")
    if not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
        os.makedirs(out_dir, exist_ok=True)
        ccbin = ["-ccbin", "/usr/bin/g++"] if os.path.exists("/usr/bin/g++") else []
        subprocess.run(["nvcc", "-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC", "-shared", *ccbin,
                        "-o", out, src], check=True, capture_output=True)
    return out
