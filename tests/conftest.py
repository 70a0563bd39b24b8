import os
import sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def pbso():
    """The product binding, with the library built in-tree first (nvcc cross-compiles on CPU)."""
    from openpbso_b200 import build
    build.build()
    import openpbso_b200
    return openpbso_b200


@pytest.fixture(scope="session")
def orc():
    from oracle import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


# Parity tolerance stated by BASELINE.json north_star for floating-point output.
REL_L2_TOL = 1e-5
MAX_ABS_TOL = 1e-6      # of full scale = max |y_ref|


def waveform_errors(y, y_ref):
    import numpy as np
    y = np.asarray(y, dtype=np.float64).ravel(); y_ref = np.asarray(y_ref, dtype=np.float64).ravel()
    full = np.max(np.abs(y_ref))
    rel_l2 = np.linalg.norm(y - y_ref) / np.linalg.norm(y_ref)
    max_abs = np.max(np.abs(y - y_ref)) / full
    return rel_l2, max_abs


def assert_waveform_parity(y, y_ref, rel=REL_L2_TOL, mx=MAX_ABS_TOL):
    rel_l2, max_abs = waveform_errors(y, y_ref)
    assert rel_l2 <= rel, "rel-L2 %.3e > %.1e" % (rel_l2, rel)
    assert max_abs <= mx, "max-abs/full-scale %.3e > %.1e" % (max_abs, mx)
    return rel_l2, max_abs
