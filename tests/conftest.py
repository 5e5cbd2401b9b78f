import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture
def fused_build():
    """select the build of the packed cluster kernels for one test (0 = 256 threads / clusters of 8, 1 = 512
    threads / clusters of 4, see qec_routing_fused_set_wide) and restore the previous choice afterwards"""
    from qecnetworks_b200 import functional as QF

    prev = []

    def select(mode):
        prev.append(QF.set_fused_wide(mode))

    yield select
    if prev:
        QF.set_fused_wide(prev[0])
