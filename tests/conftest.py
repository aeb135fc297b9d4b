import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def functor_kat():
    return np.load(os.path.join(GOLDEN, "functor_kat.npz"))


@pytest.fixture(scope="session")
def opencv_kat():
    return np.load(os.path.join(GOLDEN, "opencv_kat.npz"))


@pytest.fixture(scope="session")
def lm_kat():
    return np.load(os.path.join(GOLDEN, "lm_kat.npz"))


@pytest.fixture(scope="session")
def orc():
    from oracle import oracle

    oracle.lib()
    return oracle


def relerr(a, b, floor=1.0):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor))) if a.size else 0.0
