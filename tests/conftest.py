import os
import sys

import numpy
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box: pytest -m gpu)")
    config.addinivalue_line("markers", "slow_cpu: CPU test that takes a few seconds (full-size permutation)")


@pytest.fixture(scope="session")
def golden():
    return numpy.load(os.path.join(ROOT, "tests", "golden", "golden.npz"))


@pytest.fixture(scope="session")
def refgold():
    """Outputs of the real reference code (tests/golden/make_reference_golden.py)."""
    return numpy.load(os.path.join(ROOT, "tests", "golden", "reference_golden.npz"))


def close(a, b, rel=1e-10, abs_=1e-12):
    """BASELINE.json tolerance for indices: within 1e-10 relative OR 1e-12 absolute."""
    a, b = numpy.asarray(a, dtype=numpy.float64), numpy.asarray(b, dtype=numpy.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    d = numpy.abs(a - b)
    ok = (d <= abs_) | (d <= rel * numpy.abs(b))
    assert ok.all(), "max abs %.3e, max rel %.3e" % (d.max(), (d / numpy.maximum(numpy.abs(b), 1e-300)).max())
