"""aaerec_b200 -- B200-native AAE hot path behind the aaerec Recommender API.

Importing the package does not touch the GPU; constructing an engine/model does and fails loudly
when ``libaae_b200.so`` is missing or the device is not sm_100 (no CPU fallback).
"""
from .base import Recommender  # noqa: F401
from .condition import (ConditionList, ConditionBase, ConcatenationBasedConditioning,  # noqa: F401
                        PrecomputedEmbeddingCondition, _check_conditions)

__all__ = ["Recommender", "ConditionList", "ConditionBase", "ConcatenationBasedConditioning",
           "PrecomputedEmbeddingCondition", "AAERecommender", "AdversarialAutoEncoder", "AutoEncoder", "AAEEngine",
           "DAERecommender", "DenoisingAutoEncoder", "DecodingRecommender", "VAERecommender", "VAE"]


def __getattr__(name):
    if name in ("AAERecommender", "AdversarialAutoEncoder", "AutoEncoder"):
        from . import aae
        return getattr(aae, name)
    if name in ("DAERecommender", "DenoisingAutoEncoder"):
        from . import dae
        return getattr(dae, name)
    if name == "DecodingRecommender":
        from .decoding import DecodingRecommender
        return DecodingRecommender
    if name in ("VAERecommender", "VAE"):
        from . import vae
        return getattr(vae, name)
    if name == "AAEEngine":
        from .engine import AAEEngine
        return AAEEngine
    raise AttributeError(name)
