import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box with -m gpu)")


def golden(name):
    import numpy as np

    return np.load(os.path.join(ROOT, "tests", "golden", name), allow_pickle=False)


@pytest.fixture(scope="session")
def gold():
    return golden
