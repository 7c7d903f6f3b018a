"""The C++ facade end to end on the GPU: the re-targeted example programs must print the reference's
published output (examples/README.md:7-12 and the notebooks' captured stdout)."""
import os
import subprocess

import numpy as np
import pytest

from helpers import load_golden

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"


def _build_and_run(example, tmp_path):
    from svgdcpp_b200 import build

    lib = build.build()
    exe = tmp_path / example
    cmd = [GXX, "-std=c++17", "-O2", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", example + ".cpp"),
           "-L", os.path.dirname(lib), "-lsvgd_b200", "-Wl,-rpath," + os.path.dirname(lib), "-o", str(exe)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    run = subprocess.run([str(exe)], capture_output=True, text=True, timeout=600)
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
