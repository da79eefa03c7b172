"""``aaerec.vae`` of the overlay: the reference module's namespace with ``VAERecommender`` / ``VAE``
replaced by the B200 classes (main.py:100,122 constructs ``VAERecommender(conditions=..., **vae_params)``)."""
import importlib.util
import os
import sys

from aaerec_b200.vae import VAERecommender, VAE  # noqa: F401

from . import REFERENCE_DIR

_B200 = ("VAERecommender", "VAE")
reference_module = None
if REFERENCE_DIR is not None:
    _name = __package__ + "._reference_vae"
    _spec = importlib.util.spec_from_file_location(_name, os.path.join(REFERENCE_DIR, "vae.py"))
    try:
        _mod = importlib.util.module_from_spec(_spec)
        sys.modules[_name] = _mod
        _spec.loader.exec_module(_mod)
        reference_module = _mod
        for _k, _v in vars(_mod).items():
            if not _k.startswith("__") and _k not in _B200 and _k not in globals():
                globals()[_k] = _v
    except ImportError as _e:
        sys.modules.pop(_name, None)
        reference_import_error = _e
