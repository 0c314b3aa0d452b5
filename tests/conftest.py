"""pytest configuration: registers the ``gpu`` marker and puts the repo root on sys.path.

``-m "not gpu"`` tests run on a CPU-only container: the oracle against the golden vectors,
the host-side logic, the C-ABI symbol table, and world_size-2 gloo sharding.
``-m gpu`` tests are the parity tests proper: CUDA path (through the C ABI) vs oracle / goldens.
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        have = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture
def option():
    """Set experiment switches of libbspy_cuda.so (bspy_cuda_set_option) for one test; every switch the test touched
    returns to its default afterwards."""
    from bspy_b200 import _cuda
    touched = set()

    def set_(name, value):
        touched.add(name)
        _cuda.set_option(name, value)

    yield set_
    for name in touched:
        _cuda.set_option(name, None)
