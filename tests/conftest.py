import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden_pixels():
    return np.load(os.path.join(ROOT, "tests", "golden", "golden_pixels.npz"))


@pytest.fixture(scope="session")
def lib():
    """The product library, built in-tree (nvcc cross-compiles without a GPU)."""
    from fennec_b200 import _lib, build
    build.build()
    return _lib.load()


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle — the checker, never the thing under test."""
    from oracle import pyoracle
    pyoracle.build()
    pyoracle.set_procs(8)
    return pyoracle
