"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/aae_b200.h declares (no compute calls: there is no GPU here), and the product refuses to run
without a device instead of falling back."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "aae_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(aae_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from aaerec_b200 import _native as N
    lib = N.load()
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), "missing export: " + n
    assert set(names) == set(N.EXPORTS), set(names) ^ set(N.EXPORTS)
    assert lib.aae_version() >= 100


def test_struct_layout_matches_header():
    from aaerec_b200 import _native as N
    assert ctypes.sizeof(N.AaeDims) == 16
    assert ctypes.sizeof(N.AaeDrop) == 16
    assert ctypes.sizeof(N.StepState) == 48


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from aaerec_b200 import _native as N
    from aaerec_b200.aae import AdversarialAutoEncoder
    import scipy.sparse as sp
    import numpy as np
    m = AdversarialAutoEncoder(verbose=False)
    with pytest.raises(N.NativeError):
        m.partial_fit(sp.csr_matrix(np.eye(4, dtype=np.float32)))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "aae-recommender_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(d, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f
