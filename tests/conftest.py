import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name)))


@pytest.fixture(scope="session")
def golden_small_random():
    return load_golden("small_random.npz")


@pytest.fixture(scope="session")
def golden_small_calibrated():
    return load_golden("small_calibrated.npz")


@pytest.fixture(scope="session")
def golden_kat():
    return load_golden("kat_modules.npz")


@pytest.fixture(scope="session")
def golden_fern():
    return load_golden("fern504_subset.npz")
