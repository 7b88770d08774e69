import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def golden():
    path = os.path.join(ROOT, "tests", "golden", "golden_v1.npz")
    return dict(np.load(path, allow_pickle=False))


@pytest.fixture(scope="session")
def golden2():
    """round-2 fixtures (tests/golden/make_golden_v2.py): in-batch block at B = 256 / 512, concat block + gradient"""
    path = os.path.join(ROOT, "tests", "golden", "golden_v2.npz")
    return dict(np.load(path, allow_pickle=False))


@pytest.fixture(scope="session")
def golden3():
    """training-step fixture (tests/golden/make_golden_v3.py): the reference Discriminator through the literal D step / G step"""
    path = os.path.join(ROOT, "tests", "golden", "golden_v3.npz")
    return dict(np.load(path, allow_pickle=False))
