"""Zero-edit drop-in: an overlay of the reference's ``aaerec`` package.

Put this directory's parent (``aae-recommender_b200/``) in front of the reference on ``sys.path`` /
``PYTHONPATH`` and ``main.py``, ``eval/*.py`` and ``eval/mpd/make_submission.py`` run unchanged: every
``aaerec.<module>`` they import resolves to the reference's own file (datasets, evaluation, condition, baselines,
svd, ...), except ``aaerec.aae``, ``aaerec.dae`` and ``aaerec.vae``: those re-export the reference module's namespace
with the recommenders of main.py:98-124 replaced by the B200-native classes (``AAERecommender`` /
``AdversarialAutoEncoder`` / ``AutoEncoder`` / ``DecodingRecommender``, ``DAERecommender`` / ``DenoisingAutoEncoder``,
``VAERecommender`` / ``VAE``).  The reference's own condition objects (``aaerec.condition.ConditionList`` ...) are
accepted by the B200 classes as they are.

Where the reference lives is taken from ``AAEREC_REFERENCE`` (the directory that contains ``aaerec/``), else from
the first other ``aaerec`` package on ``sys.path``.  Without a reference package only ``aaerec.aae``,
``aaerec.condition`` and ``aaerec.base`` exist (the B200 classes).
"""
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))


def _reference_package_dir():
    cands = []
    env = os.environ.get("AAEREC_REFERENCE")
    if env:
        cands.append(os.path.join(env, "aaerec"))
    for p in sys.path:
        d = os.path.join(p or ".", "aaerec")
        if os.path.abspath(d) != _HERE:
            cands.append(d)
    for d in cands:
        if os.path.isfile(os.path.join(d, "aae.py")) and os.path.isfile(os.path.join(d, "evaluation.py")):
            return os.path.abspath(d)
    return None


REFERENCE_DIR = _reference_package_dir()
if REFERENCE_DIR is not None:
    # overlay first, reference second: aaerec.aae / .dae / .vae are ours, every other submodule is the reference's file
    __path__ = [_HERE, REFERENCE_DIR]
else:
    import aaerec_b200.base as _base
    import aaerec_b200.condition as _condition
    sys.modules[__name__ + ".base"] = _base
    sys.modules[__name__ + ".condition"] = _condition
    base, condition = _base, _condition
