import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "aae-recommender_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """gpu-marked tests need a CUDA device: without one they are skipped (instead of failing at the first), so that a
    plain ``pytest tests`` works on a CPU-only machine.  On a machine WITH a GPU nothing is skipped: a missing
    ``libaae_b200.so`` must fail loudly there (there is no fallback path to fall back to)."""
    reason = None
    try:
        import torch
        if not torch.cuda.is_available():
            reason = "no CUDA device"
    except Exception as e:   # noqa: BLE001
        reason = "torch unavailable: %r" % (e,)
    if reason is None:
        return
    skip = pytest.mark.skip(reason=reason)
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
