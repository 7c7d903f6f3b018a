"""CPU check of the error analysis behind the tensor-core pair kernel's arithmetic variants: the numpy model
of scripts/tc32_numerics_model.py (fp16 splits, fp16 kernel values, e5m2 correction terms) against the FP64 formula.  The GPU tests
measure the same quantities on the device (tests/test_gpu_tc32.py); this one keeps the analysis itself from rotting."""
import importlib.util
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _model():
    spec = importlib.util.spec_from_file_location("tc32_numerics_model", os.path.join(ROOT, "scripts", "tc32_numerics_model.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_e5m2_grid():
    m = _model()
    x = np.array([0.0, 1.0, 1.1, 1.125, 1.375, 1.9, -3.3, 2.0 ** -16, 2.0 ** -17 * 1.01, 1e6, 6.1e-5])
    q = m.e5m2(x)
    assert list(q[:7]) == [0.0, 1.0, 1.0, 1.0, 1.5, 2.0, -3.5]          # 2 mantissa bits, ties to even
    assert q[7] == 2.0 ** -16 and q[8] == 2.0 ** -16 and q[9] == 57344.0  # smallest subnormal, saturation
    assert m.e5m2(np.array([1.9]), truncate=True)[0] == 1.75             # the top byte of an fp16 value truncates
    rel = np.abs(m.e5m2(np.linspace(0.01, 100.0, 10001)) - np.linspace(0.01, 100.0, 10001)) / np.linspace(0.01, 100.0, 10001)
    assert rel.max() <= 0.125 + 1e-12


def test_fast_and_lean_variants_stay_under_the_stated_bound():
    from svgdcpp_b200 import synth

    m = _model()
    n, d = 768, 64
    x0, means, covs = synth.mvn_problem(n, d)
    X = np.ascontiguousarray(x0.T)
    G = m.mixture_grad(X, means, covs)
    fast = m.model(X, G)[0]
    lean = m.model(X, G, lean=True)[0]
    lean1 = m.model(X, G, lean=True, one_term_v=True)[0]
    print("n=%d d=%d: FAST %.3g, lean %.3g, lean with v in one term %.3g" % (n, d, fast, lean, lean1))
    assert fast < 2e-4 and lean < 2e-4
    assert lean < 2.0 * fast + 2e-5      # the e5m2 correction terms do not dominate at d = 64 ...
    assert lean1 > lean                  # ... the one-term v is the larger concession, which is why it is tied to N >= 32,768
    # d = 2: the common rounding of lo_i does not average over coordinates -- the automatic rule keeps e5m2 to d >= 48
    x0, means, covs = synth.mvn_problem(1024, 2)
    X = np.ascontiguousarray(x0.T)
    G = m.mixture_grad(X, means, covs)
    fast2, lean2 = m.model(X, G)[0], m.model(X, G, lean=True)[0]
    print("n=1024 d=2: FAST %.3g, lean %.3g" % (fast2, lean2))
    assert lean2 > fast2
