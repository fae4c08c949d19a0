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


def _n_devices():
    try:
        from geobipy_b200 import build, _lib
        build.build()
        return int(_lib.load().gbp_device_count())
    except Exception:
        return 0


def pytest_collection_modifyitems(config, items):
    """gpu-marked tests are skipped (not failed) on a box without a CUDA device, whatever -m selects."""
    gpu_items = [it for it in items if it.get_closest_marker("gpu")]
    if not gpu_items or _n_devices() > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device visible (gpu-marked test)")
    for it in gpu_items:
        it.add_marker(skip)


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
