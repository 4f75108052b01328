"""pytest configuration: the `gpu` marker and import paths.

`-m "not gpu"` covers the oracle against the reference's known-answer tables / golden vectors,
the host logic (through a NumPy test double of the backend) and the C-ABI surface of the shared
library; `-m gpu` holds the parity tests proper, which call the CUDA path through the C ABI."""

import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100a) and the built libmellon_b200.so")


def _have_gpu():
    try:
        from mellon_b200 import _native as nat

        return nat.device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device visible")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def be():
    """The product backend (one B200 through libmellon_b200.so)."""
    from mellon_b200.backend import get_backend, set_backend

    set_backend(None)
    return get_backend()
