import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def crops():
    return np.load(os.path.join(GOLDEN, "real_crops.npz"))


@pytest.fixture(scope="session")
def full1080():
    return np.load(os.path.join(GOLDEN, "real_1080p.npz"))


@pytest.fixture(scope="session")
def sweep():
    """One full-resolution pair from each of the reference's other three clips (tests/golden/make_golden_sweep.py)."""
    return np.load(os.path.join(GOLDEN, "real_sweep.npz"))


@pytest.fixture(scope="session")
def synth_small():
    return np.load(os.path.join(GOLDEN, "synth_small.npz"))


@pytest.fixture(scope="session")
def ref_funcs():
    """Outputs of the reference's OWN functions (cut out of its scripts with ast, tests/golden/make_golden_ref.py)."""
    return np.load(os.path.join(GOLDEN, "reference_funcs.npz"))


@pytest.fixture(scope="session")
def extras():
    return np.load(os.path.join(GOLDEN, "extras.npz"))


def epe(a, b):
    d = np.sqrt(((a.astype(np.float64) - b.astype(np.float64)) ** 2).sum(-1))
    return float(d.mean()), float(d.max())


def have_cv2():
    try:
        import cv2  # noqa: F401
        return True
    except Exception:
        return False
