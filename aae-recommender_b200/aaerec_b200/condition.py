"""Conditions as the AAE hot path sees them (reference: aaerec/condition.py).

The autoencoder only ever talks to its ``conditions`` object through a small protocol (aae.py:688-709,
778-780, 863-865, 944-946, 969-971): ``keys() / values() / __len__``, ``size_increment()``,
``fit_transform(raw) / transform(raw)``, ``encode(batch)``, ``zero_grad() / step()``, ``train() / eval()``.
Everything here is **duck-typed on that protocol**: the reference's own ``aaerec.condition.ConditionList``
holding ``PretrainedWordEmbeddingCondition`` / ``CategoricalCondition`` / ``EmbeddingBagCondition`` objects is
accepted as it is (``main.py:103-110`` builds exactly that), and so is this module's small stand-alone
``ConditionList`` for users without the reference package.

How a condition reaches the kernels (``CondAdapter``):
  * *row conditions* -- concatenation conditions whose ``encode`` is a parameter-free float-matrix lookup
    (``PretrainedWordEmbeddingCondition.encode`` = ``as_tensor(float32)``, condition.py:363-365; this module's
    ``PrecomputedEmbeddingCondition``): the whole [n, D] matrix is encoded once, kept in HBM and its rows are
    gathered per batch by ``aae_batch_gather`` and concatenated on the code inside ``aae_ae_fwd_bag``;
  * *generic concatenation conditions* (trainable: ``CategoricalCondition``, ``EmbeddingBagCondition``, or any other
    object whose ``impose`` concatenates along dim 1): called through their Python protocol between kernels --
    ``encode`` per batch (torch, wherever the condition keeps its parameters), rows copied into the step's
    condition buffer, and after the step the gradient of the loss w.r.t. those rows (``dec.lin1`` slice of the
    reconstruction backward, computed by the kernels) is pushed back through the condition's autograd graph
    before ``conditions.step()`` -- the reference's ``zero_grad / backward / step`` sequence (aae.py:698-709);
  * anything that does not concatenate (``ConditionalBiasing`` / ``ConditionalScaling``) would need the fused
    encoder->decoder kernel split at the code; that raises ``NotImplementedError`` (no silent CPU path).
"""
from collections import OrderedDict

import numpy as np

ROW_CONDITION_CLASS_NAMES = ("PretrainedWordEmbeddingCondition", "PrecomputedEmbeddingCondition")
_PROTOCOL = ("values", "keys", "size_increment", "encode")


def _is_condition_list(obj):
    return all(hasattr(obj, a) for a in _PROTOCOL) and hasattr(obj, "__len__")


def _check_conditions(conditions, condition_data):
    """condition.py:31-57 -- same return value and the same AssertionErrors, but any object that speaks the
    ConditionList protocol passes (the reference's class, this module's, or a user's)."""
    if not conditions and not condition_data:
        return False
    assert _is_condition_list(conditions), "`conditions` no instance of ConditionList"
    assert condition_data and conditions, "Mismatch between condition spec and supplied condition data."
    assert len(condition_data) == len(conditions), "Unexpected number of supplied condition data"
    return True


class ConditionBase(object):
    """Default behaviour of one condition (condition.py:140-255): identity fit/transform/encode, no-op optimizer
    and mode hooks.  Subclasses provide ``impose`` and ``size_increment``."""
    fusable = False   #: True when ``encode`` is a parameter-free float-matrix lookup (row condition)

    def fit(self, raw_inputs):
        return self

    def transform(self, raw_inputs):
        return raw_inputs

    def fit_transform(self, raw_inputs):
        return self.fit(raw_inputs).transform(raw_inputs)

    def encode(self, inputs):
        return inputs

    def impose(self, inputs, encoded_condition, dim=None):
        raise NotImplementedError

    def encode_impose(self, inputs, condition_input, dim=None):
        return self.impose(inputs, self.encode(condition_input), dim=None)

    def size_increment(self):
        raise NotImplementedError

    def zero_grad(self):
        return self

    def step(self):
        return self

    def train(self):
        return self

    def eval(self):
        return self


class ConcatenationBasedConditioning(ConditionBase):
    """condition.py:300-316: impose = concatenate along dim 1 (numpy or torch inputs)."""
    dim = 1

    def impose(self, inputs, encoded_condition, dim=None):
        axis = self.dim if dim is None else dim
        try:
            import torch
            if torch.is_tensor(inputs) or torch.is_tensor(encoded_condition):
                return torch.cat([torch.as_tensor(inputs), torch.as_tensor(encoded_condition).to(
                    torch.as_tensor(inputs).device)], dim=axis)
        except ImportError:   # pragma: no cover
            pass
        return np.concatenate([np.asarray(inputs), np.asarray(encoded_condition)], axis=axis)


class PrecomputedEmbeddingCondition(ConcatenationBasedConditioning):
    """Rows of a precomputed float matrix (e.g. TF-IDF-weighted word2vec title embeddings, ub.py:58-62),
    concatenated on the code -- what PretrainedWordEmbeddingCondition.encode yields once its vectoriser has run
    (condition.py:363-365)."""
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
    """Ordered ``name -> condition`` mapping speaking the protocol of condition.py:59-137; order is meaningful.
    Every list-level call fans out over the members in order."""

    def __init__(self, items):
        super(ConditionList, self).__init__(items)
        for name, c in self.items():
            assert all(hasattr(c, a) for a in ("encode", "impose", "size_increment")), \
                "condition %r does not implement the condition protocol" % (name,)

    def _each(self, method, inputs=None):
        if inputs is None:
            return [getattr(c, method)() for c in self.values() if hasattr(c, method)]
        assert len(inputs) == len(self)
        return [getattr(c, method)(x) for c, x in zip(self.values(), inputs)]

    def fit(self, raw_inputs):
        self._each("fit", raw_inputs)
        return self

    def transform(self, raw_inputs):
        return self._each("transform", raw_inputs)

    def fit_transform(self, raw_inputs):
        return self._each("fit_transform", raw_inputs)

    def encode(self, condition_inputs):
        return self._each("encode", condition_inputs)

    def encode_impose(self, x, condition_inputs, dim=None):
        assert len(condition_inputs) == len(self)
        for c, ci in zip(self.values(), condition_inputs):
            x = c.encode_impose(x, ci, dim)
        return x

    def size_increment(self):
        return sum(self._each("size_increment"))

    def zero_grad(self):
        self._each("zero_grad")
        return self

    def step(self):
        self._each("step")
        return self

    def train(self):
        self._each("train")

    def eval(self):
        self._each("eval")


def _concatenates(cond):
    """Behavioural probe: does ``cond.impose`` concatenate along dim 1?  (True for every subclass of the reference's
    ConcatenationBasedConditioning, whatever package it was imported from.)"""
    import torch
    try:
        out = cond.impose(torch.zeros(2, 3), torch.ones(2, 2))
    except Exception:   # noqa: BLE001
        return False
    return torch.is_tensor(out) and tuple(out.shape) == (2, 5) and bool((out[:, :3] == 0).all()) and \
        bool((out[:, 3:] == 1).all())


class CondAdapter(object):
    """Feeds a ConditionList-like object to the engine (see the module docstring)."""

    def __init__(self, conditions):
        self.conditions = conditions
        self.members = list(conditions.values())
        self.names = list(conditions.keys())
        self.kinds = []
        for name, c in zip(self.names, self.members):
            rowlike = bool(getattr(c, "fusable", False)) or type(c).__name__ in ROW_CONDITION_CLASS_NAMES
            if rowlike:
                self.kinds.append("rows")
            elif _concatenates(c):
                self.kinds.append("generic")
            else:
                raise NotImplementedError(
                    "condition %r (%s) does not concatenate on the code: conditional biasing/scaling would need the "
                    "fused encoder->decoder kernel split at the code; outside the accelerated envelope (no CPU "
                    "fallback)" % (name, type(c).__name__))
        self.all_rows = all(k == "rows" for k in self.kinds)
        self.size = int(conditions.size_increment())

    @staticmethod
    def _rows(encoded):
        try:
            import torch
            if torch.is_tensor(encoded):
                encoded = encoded.detach().cpu().numpy()
        except ImportError:   # pragma: no cover
            pass
        if hasattr(encoded, "toarray"):
            encoded = encoded.toarray()
        out = np.ascontiguousarray(np.asarray(encoded), dtype=np.float32)
        assert out.ndim == 2, "encoded condition must be [n, dim]"
        return out

    def encode_all_rows(self, condition_data):
        """Row conditions only: the whole [n, size_increment] float32 matrix, encoded once."""
        assert self.all_rows
        parts = [self._rows(c.encode(d)) for c, d in zip(self.members, condition_data)]
        out = parts[0] if len(parts) == 1 else np.concatenate(parts, axis=1)
        assert out.shape[1] == self.size, "size_increment() disagrees with the encoded condition width"
        return out

    @staticmethod
    def take(container, rows):
        """``container[rows]`` for the per-row containers conditions hand out (ndarray, scipy sparse, list)."""
        if isinstance(container, (list, tuple)):
            return [container[int(i)] for i in rows]
        return container[rows]

    def encode_batch(self, cond_batch, device, want_grad):
        """Generic path, one batch: returns (rows float32 [B, size] on ``device`` (detached), leaves) where leaves
        are the encoded tensors that require grad (to be given the kernels' gradient after the step)."""
        import torch
        outs, leaves, col = [], [], 0
        for c, d in zip(self.members, cond_batch):
            e = c.encode(d)
            if not torch.is_tensor(e):
                e = torch.as_tensor(self._rows(e))
            w = int(e.shape[1])
            if want_grad and e.requires_grad:
                leaves.append((e, col, col + w))
            outs.append(e.detach().to(device=device, dtype=torch.float32))
            col += w
        assert col == self.size, "size_increment() disagrees with the encoded condition width"
        rows = outs[0] if len(outs) == 1 else torch.cat(outs, dim=1)
        return rows.contiguous(), leaves

    def backward_and_step(self, leaves, grad_rows):
        """The reference's ``loss.backward()`` restricted to the conditions, then ``conditions.step()``
        (aae.py:703-709); ``conditions.zero_grad()`` ran before the batch was encoded."""
        import torch
        if leaves:
            tensors = [e for e, _, _ in leaves]
            grads = [grad_rows[:, a:b].to(device=e.device, dtype=e.dtype) for e, a, b in leaves]
            torch.autograd.backward(tensors, grads)
        self.conditions.step()
