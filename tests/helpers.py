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
