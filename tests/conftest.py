import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # `-m gpu` on a box without a GPU should skip loudly rather than fail in cudaGetDeviceCount
    if os.environ.get("SAYAL_FORCE_GPU_TESTS") == "1":
        return
    have = None
    for item in items:
        if "gpu" in item.keywords:
            if have is None:
                have = _has_gpu()
            if not have:
                item.add_marker(pytest.mark.skip(reason="no CUDA device"))


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure both shared libraries exist (the driver runs build() first; this covers bare pytest)."""
    from opensayal_b200 import LIB_PATH
    from oracle import oracle as O
    if not LIB_PATH.exists() or not O.ORACLE_LIB.exists():
        import __graft_entry__
        __graft_entry__.build()
