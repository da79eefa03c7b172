"""TEST INFRASTRUCTURE ONLY -- loader for the *unmodified* reference package.

Imports ``aaerec`` from ``/root/reference`` (read-only, authoring container) or
from the unmodified copy ``oracle/_ref`` that ``oracle/build_ref.py`` makes
(git-ignored, ships to the GPU box) so that (a) ``oracle/make_golden.py`` can
pin the restatement in ``oracle/aae_oracle.py`` against the real reference and
dump golden vectors into ``tests/golden/``, (b) ``bench.py --impl reference`` /
the ``gpu_baseline`` leg can time the reference's own code, (c) the harness
parity test can run the reference's ``Evaluation`` on both recommenders.

Nothing in the product package may import this module.

Shims (none of them touches the hot path):
  * ``gensim.models.keyedvectors.KeyedVectors`` -- imported at
    ``aaerec/aae.py:22`` and ``aaerec/ub.py``; gensim is not installed.
  * ``docutils.nodes.inline`` -- stray import at ``aaerec/condition.py:3``.
  * ``aaerec.aae.USE_WANDB = False`` -- wandb is installed, so
    ``partial_fit`` (``aaerec/aae.py:763-765``) would call ``wandb.log``
    without an initialised run.
"""
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))


def _pick_root():
    """The reference tree: AAE_REFERENCE_ROOT, else /root/reference (authoring container), else the unmodified copy
    that oracle/build_ref.py made under oracle/_ref (the GPU box)."""
    for cand in (os.environ.get("AAE_REFERENCE_ROOT"), "/root/reference", os.path.join(_HERE, "_ref")):
        if cand and os.path.isdir(os.path.join(cand, "aaerec")):
            return cand
    return "/root/reference"


REFERENCE_ROOT = _pick_root()


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "aaerec"))


def _stub(name, **attrs):
    mod = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(mod, k, v)
    sys.modules[name] = mod
    return mod


def load_reference():
    """Return the reference's ``aaerec`` package (modules aae, condition,
    evaluation imported), with the shims above applied."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    if "gensim" not in sys.modules:
        try:
            import gensim  # noqa: F401
        except ImportError:
            class KeyedVectors(object):  # duck-typed placeholder
                pass
            g = _stub("gensim")
            gm = _stub("gensim.models")
            gk = _stub("gensim.models.keyedvectors", KeyedVectors=KeyedVectors)
            g.models = gm
            gm.keyedvectors = gk
            gm.KeyedVectors = KeyedVectors
    if "docutils" not in sys.modules:
        try:
            import docutils.nodes  # noqa: F401
        except ImportError:
            d = _stub("docutils")
            d.nodes = _stub("docutils.nodes", inline=object)
    import numpy as np
    if not hasattr(np, "product"):
        np.product = np.prod  # aaerec/datasets.py:485
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import aaerec  # noqa: F401
    import aaerec.aae
    import aaerec.condition
    import aaerec.evaluation
    aaerec.aae.USE_WANDB = False
    aaerec.evaluation.wandb_is_available = False
    return aaerec
