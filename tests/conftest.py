import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def lib_built():
    """Builds libyolo_b200.so if it is missing or stale (nvcc cross-compiles without a GPU)."""
    from tensorflow_yolo_b200 import build
    return build.build()
