import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# the oracle's LAPACK/SuperLU calls are tiny; oversubscribed BLAS threads make them 10x slower
os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
os.environ.setdefault("OMP_NUM_THREADS", "1")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(scope="session")
def dre():
    """The product package (directory name contains a dot, so it is loaded through the shim)."""
    import dre_b200

    return dre_b200
