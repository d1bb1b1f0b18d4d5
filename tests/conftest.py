import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


@pytest.fixture(scope="session", autouse=True)
def _native_libs_built():
    """The suites exercise the in-tree native libraries; build them if a fresh checkout has none."""
    from molann_b200 import build
    if not (os.path.isfile(build.LIB_KERNELS) and os.path.isfile(build.LIB_TORCH)):
        build.build_all()
    yield
