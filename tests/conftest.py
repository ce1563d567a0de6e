import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    return dict(np.load(os.path.join(GOLDEN_DIR, "reference_vectors.npz")))


@pytest.fixture(scope="session")
def toy_conf():
    return os.path.join(GOLDEN_DIR, "toy", "toy.conf")


@pytest.fixture(scope="session")
def oracle():
    from oracle.oracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def reference():
    """The compiled reference, or None where oracle/_ref was not built."""
    from oracle.oracle import Reference, ref_available
    return Reference() if ref_available() else None
