import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box with -m gpu)")


def _cuda_device_present() -> bool:
    try:
        import torch

        return bool(torch.cuda.is_available())
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """`gpu` tests are SKIPPED (not failed) on a box without a CUDA device, so that a regression in the CPU-side tests is
    not buried under 'no CUDA device available' failures when somebody runs plain `pytest tests`."""
    if _cuda_device_present():
        return
    skip = pytest.mark.skip(reason="no CUDA device on this box (the hot path has no CPU fallback)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
