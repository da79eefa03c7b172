"""GPU tests of the ranking tail (masked top-k) and of the dense predict path."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _topk(scores, X, k, v_begin=0):
    from aaerec_b200 import _native as N
    B, V = scores.shape
    idx = torch.empty(B, k, dtype=torch.int32, device="cuda")
    val = torch.empty(B, k, dtype=torch.float32, device="cuda")
    ip = torch.as_tensor(X.indptr.astype(np.int32), device="cuda")
    ii = torch.as_tensor(X.indices.astype(np.int32), device="cuda")
    s = scores.clone()
    N.call("aae_masked_topk", N.ptr(s), s.stride(0), B, V, v_begin, N.ptr(ip), N.ptr(ii), k, N.ptr(idx), N.ptr(val),
           None, None)
    torch.cuda.synchronize()
    return idx.cpu().numpy(), val.cpu().numpy()


@pytest.mark.parametrize("V,k", [(64, 5), (4587, 20), (9000, 100), (200000, 100), (2000000, 100), (2000000, 500)])
def test_masked_topk_matches_reference_chain(V, k):
    """idx == argtopk(remove_non_missing(Y, X), k) (evaluation.py:183-199, 20-58) on continuous scores."""
    from aaerec_b200.synth import synth_sets
    from oracle import aae_oracle as O
    B = 6 if V > 100000 else 16
    g = torch.Generator(device="cuda").manual_seed(V + k)
    scores = torch.randn(B, V, device="cuda", generator=g)
    X = synth_sets(B, V, 12, seed=V % 97)
    idx, val = _topk(scores, X, k)
    Y = scores.cpu().numpy()
    ref = O.rank_topk(Y, X, k)
    np.testing.assert_array_equal(idx, ref)
    rows = np.arange(B)[:, None]
    np.testing.assert_array_equal(val, Y[rows, idx])
    assert not X.toarray()[rows, idx].any(), "a known item was recommended"


def test_topk_ties_and_degenerate_rows():
    """Constant rows and huge tie groups (the fallback path): the score sequence must still be the k best."""
    from aaerec_b200.synth import synth_sets
    B, V, k = 4, 50000, 50
    scores = torch.zeros(B, V, device="cuda")
    scores[1, ::7] = 1.0                       # 7143 ties at the top
    scores[2] = torch.arange(V, device="cuda") % 11
    scores[3, 123] = 5.0
    X = synth_sets(B, V, 6, seed=4)
    idx, val = _topk(scores, X, k)
    Y = scores.cpu().numpy().copy()
    Y[X.nonzero()] = -np.inf
    want = -np.sort(-Y, axis=1)[:, :k]
    np.testing.assert_array_equal(val, want)
    for b in range(B):
        assert len(set(idx[b].tolist())) == k
        np.testing.assert_array_equal(Y[b, idx[b]], want[b])


def test_topk_merge_and_shard_offsets():
    """Per-shard top-k with global ids + k-way merge == top-k over the whole vocabulary."""
    from aaerec_b200 import _native as N
    from aaerec_b200.synth import synth_sets
    B, V, k, shards = 8, 30000, 40, 3
    scores = torch.randn(B, V, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
    X = synth_sets(B, V, 10, seed=9)
    full, _ = _topk(scores, X, k)
    per = (V + shards - 1) // shards
    cv, ci = [], []
    for r in range(shards):
        lo, hi = r * per, min(V, (r + 1) * per)
        i, v = _topk(scores[:, lo:hi].contiguous(), X, k, v_begin=lo)
        cv.append(torch.as_tensor(v))
        ci.append(torch.as_tensor(i))
    cv = torch.cat(cv, 1).cuda().contiguous()
    ci = torch.cat(ci, 1).cuda().contiguous()
    oi = torch.empty(B, k, dtype=torch.int32, device="cuda")
    ov = torch.empty(B, k, dtype=torch.float32, device="cuda")
    N.call("aae_topk_merge", N.ptr(cv), N.ptr(ci), B, cv.shape[1], k, N.ptr(oi), N.ptr(ov), None)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(oi.cpu().numpy(), full)


def test_bag_fwd_matches_dense_first_layer():
    """K1 against F.normalize(X,1) @ W1^T + b1 (aae.py:132-135), including an empty row."""
    from aaerec_b200 import _native as N
    from aaerec_b200.synth import synth_sets
    V, H, B = 5000, 100, 64
    X = synth_sets(B, V, 20, seed=8).tolil()
    X[5, :] = 0
    X = X.tocsr()
    X.eliminate_zeros()
    W1 = torch.randn(H, V) * 0.1
    b1 = torch.randn(H) * 0.1
    Xd = torch.as_tensor(X.toarray(), dtype=torch.float32)
    want = torch.nn.functional.linear(torch.nn.functional.normalize(Xd, 1), W1, b1)
    W1t = W1.t().contiguous().cuda()
    out = torch.empty(B, H, device="cuda")
    ip = torch.as_tensor(X.indptr.astype(np.int32), device="cuda")
    ii = torch.as_tensor(X.indices.astype(np.int32), device="cuda")
    N.call("aae_bag_fwd", N.ptr(ip), N.ptr(ii), B, N.ptr(W1t), N.ptr(b1.cuda()), H, 1, 0, V, 1, N.ptr(out), None)
    torch.cuda.synchronize()
    np.testing.assert_allclose(out.cpu().numpy(), want.numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(out[5].cpu().numpy(), b1.numpy(), rtol=0, atol=0)
