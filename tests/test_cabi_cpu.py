"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol
include/svgd_b200.h declares, the C++ facade compiles warning-free against it, and the product path
fails loudly without a GPU (no CPU fallback).  No compute calls are made here."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "svgd_b200.h")
GXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"


@pytest.fixture(scope="module")
def built_lib():
    from svgdcpp_b200 import build

    return build.build()


def _declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(svgdb_[a-z0-9_]+)\s*\(", text)) - {"svgdb_grad_fn"})


def test_library_exports_every_declared_symbol(built_lib):
    lib = ctypes.CDLL(built_lib)
    names = _declared_symbols()
    assert len(names) >= 25
    for name in names:
        assert hasattr(lib, name), "libsvgd_b200.so does not export %s" % name


def test_python_binding_covers_the_header(built_lib):
    from svgdcpp_b200 import _capi

    assert sorted(_capi.SIGNATURES) == _declared_symbols()
    assert _capi.load().svgdb_version().startswith(b"svgd_b200")


def test_no_cpu_fallback(built_lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import numpy as np

    import svgdcpp_b200 as sv

    x0 = np.zeros((2, 10), order="F")
    model = sv.MultivariateNormal([0.0, 0.0], np.eye(2))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        sv.SVGD(2, 10, x0, sv.GaussianRBFKernel(x0), model, sv.AdaGrad(2, 10, 0.1))


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "svgdcpp_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle_binding" not in text and "svgd_oracle" not in text, f
    for dirpath, _, files in os.walk(os.path.join(ROOT, "include")):
        for f in files:
            text = open(os.path.join(dirpath, f)).read()
            assert "svgd_oracle" not in text, f


@pytest.mark.parametrize("example", ["examples/mvn_example", "examples/gmm_example", "tests/cpp/log_matrices"])
def test_facade_examples_compile(built_lib, example, tmp_path):
    """The programs for the reference's two example scenarios, written against include/SVGDCpp, and the test program of the logging / point-evaluation
    API build with the reference's own warning flags (-Wall -Wextra -Wpedantic, reference CMakeLists.txt:4)."""
    exe = tmp_path / os.path.basename(example)
    cmd = [GXX, "-std=c++17", "-Wall", "-Wextra", "-Wpedantic", "-Werror", "-O1", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, example + ".cpp"), "-L", os.path.dirname(built_lib), "-lsvgd_b200",
           "-Wl,-rpath," + os.path.dirname(built_lib), "-o", str(exe)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr


def test_mini_eigen_prints_like_eigen(built_lib, tmp_path):
    """`3 * Eigen::MatrixXd::Random(2, 10)` printed with the stand-in matrix type reproduces the
    'Initial particle coordinates' block of reference examples/README.md:7-9 character for character."""
    src = tmp_path / "p.cpp"
    src.write_text('#include <iostream>\n#include "SVGDCpp/MiniEigen.hpp"\n'
                   "int main(){ Eigen::MatrixXd m = 3 * Eigen::MatrixXd::Random(2, 10); std::cout << m << std::endl; }\n")
    exe = tmp_path / "p"
    res = subprocess.run([GXX, "-std=c++17", "-DSVGDCPP_NO_EIGEN", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)],
                         capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    out = subprocess.run([str(exe)], capture_output=True, text=True).stdout
    expected = ("  2.04113    1.6986   2.46988 -0.988663  -1.33335 -0.135618 -0.811293   2.71338   0.81427  -2.15038\n"
                "-0.633702   1.79064  -1.81469   1.60938   0.32382  0.773226 0.0804055   2.49717   1.30378  0.641813\n")
    assert out == expected


def test_bracket_predictor_order_on_recorded_trajectory():
    """The median bracket is predicted by cubic extrapolation of the last four medians (median_scale in svgd_b200_api.cu).
    On the trajectory recorded at the bench shape the cubic predictor must beat the quadratic one by a clear margin in the Adam
    transient and stay below the smallest bracket half-width (2e-5) in the stationary phase."""
    import json
    import os

    import numpy as np

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "median_trajectory.json")
    m = np.array(json.load(open(path))["med2"])
    quad = np.abs(3 * m[2:-1] - 3 * m[1:-2] + m[:-3] - m[3:]) / m[3:]              # predicts m[k] from m[k-1..k-3]
    cub = np.abs(4 * m[3:-1] - 6 * m[2:-2] + 4 * m[1:-3] - m[:-4] - m[4:]) / m[4:]   # ... from m[k-1..k-4]
    transient_q, transient_c = np.median(quad[2:22]), np.median(cub[1:21])            # steps 5..25
    assert transient_c < 0.3 * transient_q, (transient_q, transient_c)
    assert np.median(cub[60:]) < 2e-5


def test_eigen_text_format_of_the_python_mirror():
    """The Python mirror prints matrices the way Eigen's operator<< does (6 significant digits, right-aligned to the widest
    entry): the published stdout block of the reference's mvn example (examples/README.md:7-12) is the known answer."""
    from helpers import load_golden
    from svgdcpp_b200.svgd import _eigen_str

    g = load_golden("mvn_example")
    init = np.array(g["initial"]).T
    assert _eigen_str(init) == ("  2.04113    1.6986   2.46988 -0.988663  -1.33335 -0.135618 -0.811293   2.71338   0.81427  -2.15038\n"
                                "-0.633702   1.79064  -1.81469   1.60938   0.32382  0.773226 0.0804055   2.49717   1.30378  0.641813")


def test_header_is_plain_c(tmp_path):
    """include/svgd_b200.h is a C ABI: it must compile as C99 with pedantic warnings as errors (no C++ or torch types leak in)."""
    src = tmp_path / "abi.c"
    src.write_text('#include "svgd_b200.h"\nint main(void) { return svgdb_version() == 0; }\n')
    res = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), str(src)],
                         capture_output=True, text=True)
    assert res.returncode == 0, res.stderr


def test_host_math_unit_tests(tmp_path):
    """The library's pure host arithmetic (svgdcpp_b200/csrc/host_math.hpp: median keys, the Cholesky factor of the Hessian scale,
    the row chunks of svgdb_step_host) is included verbatim by svgd_b200_api.cu; its unit test builds and runs on the CPU."""
    exe = tmp_path / "host_math_test"
    res = subprocess.run([GXX, "-std=c++17", "-O1", "-Wall", "-Wextra", "-Wpedantic", "-Werror", "-I", os.path.join(ROOT, "svgdcpp_b200", "csrc"),
                          os.path.join(ROOT, "tests", "cpp", "host_math_test.cpp"), "-o", str(exe)], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    run = subprocess.run([str(exe)], capture_output=True, text=True)
    assert run.returncode == 0 and "all checks passed" in run.stdout, run.stdout
    assert '#include "host_math.hpp"' in open(os.path.join(ROOT, "svgdcpp_b200", "csrc", "svgd_b200_api.cu")).read()
