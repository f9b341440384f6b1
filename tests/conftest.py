import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch

        have_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        have_gpu = False
    if have_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def pytest_sessionfinish(session, exitstatus):
    """Write the fractions the parity comparisons observed (tests/helpers.py OBSERVED) next to the required ones."""
    try:
        import json

        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        from helpers import OBSERVED
    except Exception:  # pragma: no cover
        return
    if not OBSERVED:
        return
    out_dir = os.path.join(ROOT, "gpurun_out")
    if not os.path.isdir(out_dir):
        return
    rows = {k: {**v, "margin": v["observed_min"] - v["required"]} for k, v in sorted(OBSERVED.items())}
    with open(os.path.join(out_dir, "parity_observed.json"), "w") as f:
        json.dump(rows, f, indent=1)
