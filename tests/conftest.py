import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden(name):
    return np.load(os.path.join(GOLDEN_DIR, name + ".npz"))


def rel_to_max(a, b):
    """max |a-b| / max |b| — the 'relative tolerance' all value checks use (stated per test)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.fixture(scope="session")
def geo():
    return golden("geometry")


@pytest.fixture(scope="session")
def gold_pyramids():
    v = golden("volume")
    return [v[f"a{l}"] for l in range(4)], [v[f"b{l}"] for l in range(4)]
