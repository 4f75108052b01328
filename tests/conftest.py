"""pytest configuration: the `gpu` marker and import paths.

`-m "not gpu"` covers the oracle against the reference's known-answer tables / golden vectors,
the host logic (through a NumPy test double of the backend) and the C-ABI surface of the shared
library; `-m gpu` holds the parity tests proper, which call the CUDA path through the C ABI."""

import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100a) and the built libmellon_b200.so")
    config.addinivalue_line("markers", "run_last: written after the round's GPU minutes were spent — green on the NumPy test "
                                       "double of the ABI, not yet run on hardware; ordered after every test that has been, "
                                       "so that `-x` cannot hide those behind a first-run surprise")


def _have_gpu():
    """True when the CUDA library sees a device.  A box that HAS a GPU but cannot load the library (stale build,
    missing symbol) must not turn the GPU suite into 385 silent skips: that is an error, raised here."""
    gpu_node = os.path.exists("/dev/nvidia0")
    try:
        from mellon_b200 import _native as nat

        return nat.device_count() > 0
    except Exception as e:
        if gpu_node:
            raise pytest.UsageError(f"a GPU is present but libmellon_b200.so cannot be used: {e}")
        return False


def pytest_collection_modifyitems(config, items):
    items.sort(key=lambda item: 1 if "run_last" in item.keywords else 0)      # stable: file order otherwise kept
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device visible")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


_cuda_backend = None


@pytest.fixture(params=["fake", pytest.param("cuda", marks=pytest.mark.gpu)])
def be(request):
    """The backend under test.

    ``cuda`` (marked ``gpu``): the product backend — one B200 through libmellon_b200.so.
    ``fake``: the NumPy test double of the C ABI (tests/fake_lib.py), which runs the same host
    logic (estimators, validation, program compilation, ctypes marshalling) on a box without a
    GPU; the parity assertions then check the double and the host side against the oracle."""
    global _cuda_backend
    from mellon_b200.backend import CudaBackend, set_backend

    if request.param == "cuda":
        if _cuda_backend is None:
            _cuda_backend = CudaBackend.from_environment()
        backend = _cuda_backend
    else:
        from fake_lib import FakeBackend

        backend = FakeBackend()
    set_backend(backend)
    yield backend
    set_backend(None)
