"""pytest configuration: registers the `gpu` marker and puts the repo root on sys.path.

`-m "not gpu"` covers the oracle (against the reference's golden vectors and KATs), the host
logic, the build-time generator and the C-ABI symbol table -- no compute call is made without a
GPU.  `-m gpu` tests are the parity tests proper and call through the C ABI on cuda:0.
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _cuda_present() -> bool:
    return os.path.exists("/dev/nvidiactl") or os.path.exists("/dev/nvidia0")


def pytest_collection_modifyitems(config, items):
    if _cuda_present():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    return np.load(os.path.join(ROOT, "tests", "golden", "ref_torch_golden.npz"))
