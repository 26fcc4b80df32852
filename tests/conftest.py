import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def built_lib():
    """the in-tree CUDA extension; built on demand (nvcc cross-compiles without a GPU)"""
    from orb_slam2_aruco_b200 import build
    build.build()
    from orb_slam2_aruco_b200 import _lib
    return _lib.lib()
