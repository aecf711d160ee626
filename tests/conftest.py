"""pytest configuration.

`-m "not gpu"`: oracle vs golden vectors, host logic, C-ABI surface.  `-m gpu`: parity of the
CUDA path (through the C ABI) against the oracle on a real B200.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build(ref=True)
    return O


@pytest.fixture(scope="session")
def lib():
    """The product library, built in-tree.  nvcc cross-compiles without a GPU."""
    from rminc_b200 import build
    build.build()
    from rminc_b200 import _lib
    return _lib.load()


def have_gpu() -> bool:
    try:
        from rminc_b200 import _lib
        return _lib.load().rmg_device_count() > 0
    except Exception:
        return False
