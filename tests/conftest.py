import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def hexagon_p():
    """C-6 fixture: sixth_scenario.py:291-292,308-310 (unit hexagon, antipodal swap)."""
    import numpy as np
    s3 = np.sqrt(3) / 2
    start = np.array([[s3, 0.5, -2.618], [0, 1, -1.571], [-s3, 0.5, -0.524],
                      [-s3, -0.5, 0.524], [0, -1, 1.571], [s3, -0.5, 2.618]])
    goal = -start.copy()
    goal[:, 2] = start[:, 2]
    return np.concatenate([start.ravel(), goal.ravel()])


@pytest.fixture(scope="session")
def pkg():
    """The product package (ctypes over libnmpc_b200.so), imported under the alias nmpc_b200."""
    import __graft_entry__ as ge
    return ge.load_package()
