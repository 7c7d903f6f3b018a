"""The C++ facade end to end on the GPU: the programs for the reference's example scenarios must print the reference's
published output (examples/README.md:7-12 and the notebooks' captured stdout)."""
import os
import subprocess

import numpy as np
import pytest

from helpers import load_golden

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"


def _build_and_run(example, tmp_path, src_dir="examples", args=()):
    from svgdcpp_b200 import build

    lib = build.build()
    exe = tmp_path / example
    cmd = [GXX, "-std=c++17", "-O2", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, src_dir, example + ".cpp"),
           "-L", os.path.dirname(lib), "-lsvgd_b200", "-Wl,-rpath," + os.path.dirname(lib), "-o", str(exe)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    run = subprocess.run([str(exe), *args], capture_output=True, text=True, timeout=600)
    assert run.returncode == 0, run.stdout + run.stderr
    return run.stdout


def _blocks(stdout):
    lines = stdout.strip().splitlines()
    i0, i1 = lines.index("Initial particle coordinates"), lines.index("Final particle coordinates")
    parse = lambda ls: np.array([[float(t) for t in l.split()] for l in ls])
    return parse(lines[i0 + 1:i1]), parse(lines[i1 + 1:i1 + 3])


@pytest.mark.parametrize("example", ["mvn_example", "gmm_example"])
def test_cpp_examples_print_reference_output(example, tmp_path):
    g = load_golden(example)
    out = _build_and_run(example, tmp_path)
    init, final = _blocks(out)
    assert np.array_equal(init.T, np.array(g["initial"]))   # printed digits, exactly
    assert np.array_equal(final.T, np.array(g["final"]))


def test_cpp_mvn_example_stdout_is_the_readme_block(tmp_path):
    out = _build_and_run("mvn_example", tmp_path)
    expected = """Initial particle coordinates
  2.04113    1.6986   2.46988 -0.988663  -1.33335 -0.135618 -0.811293   2.71338   0.81427  -2.15038
-0.633702   1.79064  -1.81469   1.60938   0.32382  0.773226 0.0804055   2.49717   1.30378  0.641813
Final particle coordinates
 0.469815 -0.184629 0.0827075  -1.04192 -0.946601  -1.73173  -1.14872  0.452507 -0.791678  -2.05712
  1.16686   1.82829 -0.375293   2.64404 -0.148336   1.15547  -1.79038   3.21318  0.828686 -0.556122
"""
    assert out == expected


def _parse_log(text):
    """The reference's intermediate-matrices log (SVGD.hpp:346-359) -> [{LogModelGrad, Kernel, KernelGrad, CoordMat}] per step."""
    steps = []
    for k, block in enumerate(text.split("========== Step ")[1:]):
        head, _, rest = block.partition(" ==========\n")
        assert int(head) == k + 1
        mats = {}
        for name in ("LogModelGrad", "Kernel", "KernelGrad", "CoordMat"):
            assert rest.startswith(name + "=\n"), rest[:40]
            body, _, rest = rest[len(name) + 2:].partition("\n\n")
            mats[name] = np.array([[float(t) for t in line.split()] for line in body.splitlines()])
        assert rest == ""
        steps.append(mats)
    return steps


def test_cpp_log_intermediate_matrices(tmp_path, oracle):
    """SVGDOptions::LogIntermediateMatrices through the C++ facade: the file has the reference's layout (SVGD.hpp:346-359) and every
    printed number (Eigen's 6 significant digits) is the oracle's: grad log p, K(j, i) = k(x_j, x_i), the (n d) x n kernel-gradient
    matrix and the updated coordinates of each step."""
    from helpers import assert_matches_printed

    log = tmp_path / "log.txt"
    stdout = _build_and_run("log_matrices", tmp_path, src_dir=os.path.join("tests", "cpp"), args=[str(log)])
    steps = _parse_log(log.read_text())
    n, d, iters = 6, 2, 3
    assert len(steps) == iters
    mean, cov = np.array([[0.5, -0.25]]), np.array([[[0.5, 0.2], [0.2, 0.8]]])
    X0 = np.array([[1.5, -0.75, 0.25, 2.0, -1.25, 0.5], [-0.5, 1.0, 0.75, -1.5, 0.125, 2.25]]).T.copy()
    X = X0
    for t, m in enumerate(steps):
        assert m["LogModelGrad"].shape == (d, n) and m["Kernel"].shape == (n, n)
        assert m["KernelGrad"].shape == (n * d, n) and m["CoordMat"].shape == (d, n)
        a = oracle.rbf_median_scale(X)
        K, dK = oracle.kernel_matrices(X, a)                       # K[i, j] = k(x_j, x_i), dK[i, j, :] = grad k(x_j, x_i)
        assert_matches_printed(oracle.mvn_sum_logp_grad(X, mean, cov).T, m["LogModelGrad"])
        assert_matches_printed(K.T, m["Kernel"])
        assert_matches_printed(dK.reshape(n, n * d).T, m["KernelGrad"])
        X = oracle.svgd_run(X0, t + 1, mean, cov, opt_kind=oracle.OPT_ADAM, lr=0.1)
        assert_matches_printed(X.T, m["CoordMat"])
    # the model's point evaluations printed by the same program (17 digits): unnormalised and normalised Gaussian at x = (0.75, -1.5)
    vals = np.array([float(t) for t in stdout.strip().splitlines()[-1].split()])
    x = np.array([[0.75, -1.5]])
    logp = oracle.mvn_sum_logp(x, mean, cov)[0]
    g = oracle.mvn_sum_logp_grad(x, mean, cov)[0]
    norm = 1.0 / (2.0 * np.pi * np.sqrt(np.linalg.det(cov[0])))
    expect = np.array([logp, np.exp(logp), g[0], g[1], np.log(norm) + logp, norm * np.exp(logp), norm * np.exp(logp) * g[0], norm * np.exp(logp) * g[1]])
    assert np.max(np.abs(vals - expect) / np.maximum(1e-300, np.abs(expect))) < 1e-13
