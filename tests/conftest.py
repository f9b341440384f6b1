import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line(
        "markers",
        "hw_unverified: GPU test of code written after this round's GPU budget was spent - its kernel bodies are "
        "checked on the CPU by tests/emu, but it has never run on a B200; skipped unless SNRF_RUN_UNVERIFIED=1",
    )


def pytest_collection_modifyitems(config, items):
    try:
        import torch

        have_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        have_gpu = False
    if os.environ.get("SNRF_RUN_UNVERIFIED", "0") != "1":
        unverified = pytest.mark.skip(reason="not yet run on hardware (set SNRF_RUN_UNVERIFIED=1 to run it)")
        for item in items:
            if "hw_unverified" in item.keywords:
                item.add_marker(unverified)
    if have_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
