import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def fixture_unweighted():
    return np.fromfile(os.path.join(GOLDEN, "rmat10_1024.bin"), dtype="<u4").reshape(-1, 2)


@pytest.fixture(scope="session")
def fixture_weighted():
    return np.fromfile(os.path.join(GOLDEN, "rmat10_1024_w.bin"), dtype="<u4").reshape(-1, 3)


@pytest.fixture(scope="session")
def golden_fixture():
    return np.load(os.path.join(GOLDEN, "fixture.npz"))


@pytest.fixture(scope="session")
def golden_rmat12():
    return np.load(os.path.join(GOLDEN, "rmat12_seed12.npz"))


@pytest.fixture(scope="session")
def rmat12():
    from graphtap_b200.rmat import rmat_edges
    return rmat_edges(12, seed=12, weighted=True)
