import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: long CPU-emulation run, skipped unless DKTB_SLOW=1")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if os.environ.get("DKTB_SLOW", "0") != "1":
        slow = pytest.mark.skip(reason="slow CPU-emulation test (set DKTB_SLOW=1); covered on the GPU by tests/test_dkt_gpu.py")
        for item in items:
            if "slow" in item.keywords:
                item.add_marker(slow)
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
