import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    """The C oracle (test infrastructure), compiled on demand with gcc."""
    import oracle_py
    oracle_py.build()
    return oracle_py


@pytest.fixture(scope="session")
def built_lib():
    """The CUDA extension, cross-compiled on demand with nvcc (no GPU needed to build/load)."""
    from geobipy_b200 import build, _lib
    build.build()
    return _lib.load()


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
