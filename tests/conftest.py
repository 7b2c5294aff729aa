import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: long CPU replay, enabled with TRAJOPT_SLOW=1")


def _have_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    slow = os.environ.get("TRAJOPT_SLOW") == "1"
    skip_slow = pytest.mark.skip(reason="set TRAJOPT_SLOW=1 to run full-length golden replays")
    skip_gpu = None if _have_cuda() else pytest.mark.skip(reason="no CUDA device (GPU parity tests run with -m gpu on the B200 box)")
    for item in items:
        if "slow" in item.keywords and not slow:
            item.add_marker(skip_slow)
        if "gpu" in item.keywords and skip_gpu is not None:
            item.add_marker(skip_gpu)
