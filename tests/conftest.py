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
    """gpu-marked tests need a CUDA device and the built library: skip them (instead of failing at the first one)
    when either is missing, so that a plain ``pytest tests`` works on a CPU-only machine."""
    reason = None
    try:
        import torch
        if not torch.cuda.is_available():
            reason = "no CUDA device"
    except Exception as e:   # noqa: BLE001
        reason = "torch unavailable: %r" % (e,)
    if reason is None and not os.path.exists(os.path.join(PKG, "aaerec_b200", "libaae_b200.so")):
        reason = "libaae_b200.so not built"
    if reason is None:
        return
    skip = pytest.mark.skip(reason=reason)
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
