"""Condition protocol as the hot path sees it (reference: aaerec/condition.py).

The AAE only ever calls ``conditions.encode_impose(z, c_batch)``, ``.size_increment()``,
``.zero_grad()``, ``.step()``, ``.train()/.eval()``, ``.keys()``, ``.fit_transform()`` and
``.transform()`` (aae.py:688-709, 778-780, 863-865, 944-946, 969-971).  This module mirrors the classes
on that protocol: ``ConditionList`` (condition.py:59-137), ``ConditionBase`` (140-255),
``ConcatenationBasedConditioning`` (300-316) and a precomputed-embedding condition equivalent to
``PretrainedWordEmbeddingCondition`` (345-369) once its TF-IDF-weighted word vectors exist.

Only concatenation conditions whose ``encode`` is a pure float-matrix lookup are fused into the CUDA
path (the condition rows are copied to the device and concatenated on the code inside
``aae_ae_fwd``).  Trainable conditions keep their own torch modules/optimizers in the reference; they
are outside the accelerated envelope and raise ``NotImplementedError`` here.
"""
from abc import ABC, abstractmethod
from collections import OrderedDict

import numpy as np


def _check_conditions(conditions, condition_data):
    """condition.py:31-57 -- same return value and the same AssertionErrors."""
    if not conditions and not condition_data:
        return False
    assert isinstance(conditions, ConditionList), "`conditions` no instance of ConditionList"
    assert condition_data and conditions, "Mismatch between condition spec and supplied condition data."
    assert len(condition_data) == len(conditions), "Unexpected number of supplied condition data"
    return True


class ConditionBase(ABC):
    """condition.py:140-255: fit/transform on raw inputs, encode/impose on batches, optional
    optimizer callbacks (no-ops for parameter-free conditions)."""

    def fit(self, raw_inputs):
        return self

    def transform(self, raw_inputs):
        return raw_inputs

    def fit_transform(self, raw_inputs):
        return self.fit(raw_inputs).transform(raw_inputs)

    @abstractmethod
    def encode(self, inputs):
        """ batch of transformed inputs -> float rows """

    @abstractmethod
    def impose(self, inputs, encoded_condition, dim=None):
        """ combine code and encoded condition """

    def encode_impose(self, inputs, condition_input, dim=None):
        return self.impose(inputs, self.encode(condition_input), dim)

    @abstractmethod
    def size_increment(self):
        """ how much the code grows """

    def zero_grad(self):
        return self

    def step(self):
        return self

    def train(self):
        return self

    def eval(self):
        return self

    #: True when ``encode`` is a parameter-free float-matrix lookup (fusable into the kernels)
    fusable = False


class ConcatenationBasedConditioning(ConditionBase):
    """condition.py:300-316: impose = concatenate along dim 1."""
    dim = 1

    def impose(self, inputs, encoded_condition, dim=None):
        return np.concatenate([np.asarray(inputs), np.asarray(encoded_condition)], axis=self.dim if dim is None else dim)


class PrecomputedEmbeddingCondition(ConcatenationBasedConditioning):
    """Rows of a precomputed float matrix (e.g. TF-IDF-weighted word2vec title embeddings,
    ub.py:58-62), concatenated on the code -- what PretrainedWordEmbeddingCondition.encode yields
    (condition.py:363-365): ``as_tensor(inputs, float32)``."""
    fusable = True

    def __init__(self, dim):
        self._dim = int(dim)

    def encode(self, inputs):
        out = np.ascontiguousarray(np.asarray(inputs), dtype=np.float32)
        assert out.ndim == 2 and out.shape[1] == self._dim, "condition rows must be [n, %d]" % self._dim
        return out

    def size_increment(self):
        return self._dim


class ConditionList(OrderedDict):
    """condition.py:59-137: ordered name -> condition mapping; order is meaningful."""

    def __init__(self, items):
        super(ConditionList, self).__init__(items)
        assert all(isinstance(v, ConditionBase) for v in self.values())

    def fit(self, raw_inputs):
        assert len(raw_inputs) == len(self)
        for cond, cond_inp in zip(self.values(), raw_inputs):
            cond.fit(cond_inp)
        return self

    def transform(self, raw_inputs):
        assert len(raw_inputs) == len(self)
        return [c.transform(inp) for c, inp in zip(self.values(), raw_inputs)]

    def fit_transform(self, raw_inputs):
        assert len(raw_inputs) == len(self)
        return [cond.fit_transform(inp) for cond, inp in zip(self.values(), raw_inputs)]

    def encode_impose(self, x, condition_inputs, dim=None):
        assert len(condition_inputs) == len(self)
        for condition, condition_input in zip(self.values(), condition_inputs):
            x = condition.encode_impose(x, condition_input, dim)
        return x

    def encode(self, condition_inputs):
        assert len(condition_inputs) == len(self)
        return [c.encode(ci) for c, ci in zip(self.values(), condition_inputs)]

    def zero_grad(self):
        for condition in self.values():
            condition.zero_grad()
        return self

    def step(self):
        for condition in self.values():
            condition.step()
        return self

    def size_increment(self):
        return sum(v.size_increment() for v in self.values())

    def train(self):
        for condition in self.values():
            if hasattr(condition, 'train'):
                condition.train()

    def eval(self):
        for condition in self.values():
            if hasattr(condition, 'eval'):
                condition.eval()

    def fused_rows(self, condition_inputs):
        """Concatenate the encoded rows of all (fusable) conditions: float32 [B, size_increment()]."""
        for name, c in self.items():
            if not getattr(c, "fusable", False) or not isinstance(c, ConcatenationBasedConditioning):
                raise NotImplementedError(
                    "condition %r (%s) is outside the accelerated envelope: only concatenation conditions whose "
                    "encode() is a float-matrix lookup are fused; no CPU fallback" % (name, type(c).__name__))
        enc = self.encode(condition_inputs)
        return enc[0] if len(enc) == 1 else np.concatenate(enc, axis=1)
