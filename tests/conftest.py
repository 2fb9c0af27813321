import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def nm_lib():
    """The C-ABI library; built on demand in the CPU container (nvcc cross-compiles sm_100a)."""
    from neuromancer_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        from neuromancer_b200 import build
        build.build_library(build.aot_plans())
    return _lib.load()
