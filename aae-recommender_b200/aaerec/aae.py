"""``aaerec.aae`` of the overlay: the reference module's namespace with the hot-path classes replaced.

``from aaerec.aae import AAERecommender, DecodingRecommender`` (main.py:13) keeps working: ``AAERecommender``,
``AdversarialAutoEncoder``, ``AutoEncoder`` and ``DecodingRecommender`` are the CUDA-backed ones; names the B200 package
does not provide (``Encoder``, ``Decoder`` ...) come from the reference's own ``aae.py`` when it is importable.
"""
import importlib.util
import os
import sys

from aaerec_b200.aae import AAERecommender, AdversarialAutoEncoder, AutoEncoder  # noqa: F401
from aaerec_b200.decoding import DecodingRecommender  # noqa: F401

from . import REFERENCE_DIR

_B200 = ("AAERecommender", "AdversarialAutoEncoder", "AutoEncoder", "DecodingRecommender")
reference_module = None
if REFERENCE_DIR is not None:
    _name = __package__ + "._reference_aae"
    _spec = importlib.util.spec_from_file_location(_name, os.path.join(REFERENCE_DIR, "aae.py"))
    try:
        _mod = importlib.util.module_from_spec(_spec)
        sys.modules[_name] = _mod
        _spec.loader.exec_module(_mod)      # relative imports (.base, .condition ...) resolve through the overlay
        reference_module = _mod
        if hasattr(_mod, "USE_WANDB") and "WANDB_API_KEY" not in os.environ:
            _mod.USE_WANDB = False          # wandb.log without an initialised run (aae.py:763-765)
        for _k, _v in vars(_mod).items():
            if not _k.startswith("__") and _k not in _B200 and _k not in globals():
                globals()[_k] = _v
    except ImportError as _e:               # e.g. gensim missing: the B200 classes alone are still usable
        sys.modules.pop(_name, None)
        reference_import_error = _e
