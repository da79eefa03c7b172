"""Parity at the BASELINE configs the round-1 tests did not reach ('parity holes'):
  * the MPD shape end to end: V = 2,000,000, batch 100 -- two partial_fit steps and the fused top-k against the oracle;
  * whole partial_fit steps at the reference scripts' larger batches (500: main.py:76; 1000: mpd.py:75-76), whatever
    kernel serves them;
  * init_uniform's item-shard layout (the on-device init of every MPD bench leg) against the single-shard matrices;
  * the device-side epoch feed (aae_batch_gather) against scipy row indexing.
All calls go through the C ABI."""
import numpy as np
import pytest
import torch

from helpers import rel_err

pytestmark = pytest.mark.gpu

TOL = 1e-4


def _oracle_and_model(V, H, C, B, dropout, seed=42, **kw):
    from oracle import aae_oracle as O
    from aaerec_b200.aae import AdversarialAutoEncoder
    params = O.init_params(V, H, C, seed=seed)
    oracle = O.OracleAAE(params, n_code=C)
    model = AdversarialAutoEncoder(n_hidden=H, n_code=C, batch_size=B, dropout=dropout, verbose=False, rng="oracle", **kw)
    model._build(V, C, params={k: v.clone() for k, v in params.items()})
    return O, oracle, model


def _steps_vs_oracle(V, B, steps, mean_len, dropout=(.2, .2), weight_keys=None, **kw):
    from aaerec_b200.synth import synth_sets
    H, C = 100, 50
    O, oracle, model = _oracle_and_model(V, H, C, B, dropout, **kw)
    # Data seed 23: with seeds 21 / 25 / 26 one first-layer encoder unit of one row sits within fp32 rounding of its ReLU
    # kink in step 2, where the summation order of the W1 row updates decides its side (measured: 8 identical repeats
    # land 0-8 times on the oracle's side, scripts/dbg_race.py); Adam's normalisation then turns that one element into
    # a 2e-3 relative difference of enc.lin1.bias.  Both outcomes are valid fp32 evaluations of the reference; the
    # gate needs data without such a coincidence.
    X = synth_sets(B * steps, V, mean_len, seed=23)
    torch.manual_seed(13)
    for s in range(steps):
        xb = X[s * B:(s + 1) * B]
        st = torch.get_rng_state()
        model.partial_fit(xb)
        got = model.losses()
        torch.set_rng_state(st)
        want = oracle.partial_fit(xb.toarray(), None, O.draw_step_rng(B, H, C, dropout))
        np.testing.assert_allclose(got, want, rtol=TOL)
    sd = model.state_dict()
    for k, v in oracle.p.items():
        if weight_keys is None or k in weight_keys:
            assert rel_err(sd[k].numpy(), v.numpy()) < TOL, (k, rel_err(sd[k].numpy(), v.numpy()))
    return O, oracle, model, X


@pytest.mark.parametrize("impl", ["auto", "simt"])
@pytest.mark.parametrize("B", [500, 1000])
def test_partial_fit_parity_at_script_batch_sizes(B, impl):
    """main.py:76 (500) and eval/mpd/mpd.py:75-76 (1000): losses and all 18 weight tensors after 3 steps, through the
    kernel the engine picks (the chunked tcgen05 path) and through the fp32 CUDA-core kernel."""
    _, _, model, _ = _steps_vs_oracle(V=6000, B=B, steps=3, mean_len=12, impl=impl)
    assert model.engine.impl_for(B) == (1 if impl == "auto" else 0)


def test_partial_fit_parity_ragged_large_batch():
    """a batch that is neither a multiple of the row chunk nor of 8"""
    _steps_vs_oracle(V=4000, B=333, steps=2, mean_len=9)


def test_mpd_shape_train_and_fused_topk_vs_oracle():
    """BASELINE configs[3]/[4]: V = 2M, batch 100.  Two partial_fit steps (oracle RNG, dropout on) against the dense
    CPU oracle, then the fused predict_topk for 64 rows at k = 100 and k = 500 against the oracle's ranking chain."""
    from aaerec_b200.synth import synth_sets
    V, B = 2000000, 100
    O, oracle, model, X = _steps_vs_oracle(V=V, B=B, steps=2, mean_len=66)
    Xq = synth_sets(64, V, 25, 1, 100, seed=77)
    dense = Xq.toarray()
    logits = oracle.logits(dense)
    for k in (100, 500):
        top, val = model.predict_topk(Xq, k, return_scores=True)
        assert model.engine.topk_fallbacks == 0
        ref = O.rank_topk(oracle.predict(dense), dense, k)
        mism = top != ref
        rows = np.arange(64)[:, None]
        # wherever the indices differ the oracle's logits must be tied to fp32 noise
        assert np.all(np.abs(logits[rows, top][mism] - logits[rows, ref][mism]) <= 2e-6 * np.abs(logits[rows, ref][mism]) + 1e-7)
        assert mism.mean() < 0.01, mism.mean()


def test_init_uniform_shards_are_slices_of_the_single_gpu_matrices():
    """The rows an item-sharded engine draws are exactly the rows [v_begin, v_end) of what a single engine draws, for
    shard boundaries inside and on generator-block borders; the replicated small layers are identical."""
    from aaerec_b200.engine import AAEEngine
    V, H, C = 200003, 100, 50
    one = AAEEngine(V, H, C, max_batch=8)
    one.init_uniform(42)
    bound = 1.0 / np.sqrt(H)
    assert 0.99 * bound < float(one.Wd3.abs().max()) <= bound * (1 + 1e-6)
    assert abs(float(one.Wd3.mean())) < 1e-3 and float(one.W1t.abs().max()) <= (1 + 1e-6) / np.sqrt(V)
    for world in (2, 3, 8):
        for rank in range(world):
            sh = AAEEngine(V, H, C, max_batch=8, rank=rank, world=world, exchange="none")
            sh.init_uniform(42)
            lo, hi = sh.v_begin, sh.v_end
            assert torch.equal(sh.Wd3[: hi - lo], one.Wd3[lo:hi]) and torch.equal(sh.W1t[: hi - lo], one.W1t[lo:hi])
            assert torch.equal(sh.bd3[: hi - lo], one.bd3[lo:hi])
            assert torch.equal(sh.enc, one.enc) and torch.equal(sh.dec, one.dec) and torch.equal(sh.disc, one.disc)
            del sh


def test_batch_gather_matches_scipy_row_indexing():
    """aae_batch_gather (device-side shuffle + batching, aae.py:815-823) == X[perm][start:end] and cond[perm][start:end]."""
    from aaerec_b200.engine import AAEEngine
    from aaerec_b200.synth import synth_sets, synth_condition
    n, V, D = 2357, 5000, 12
    X = synth_sets(n, V, 7, min_len=0, seed=4)
    X = X.tolil()
    X[5] = 0                                              # an empty row
    X = X.tocsr()
    X.eliminate_zeros()
    cond = synth_condition(n, D)
    eng = AAEEngine(V, 100, 50, cond_dim=D, max_batch=64)
    eng.set_epoch_data(X.indptr, X.indices, cond)
    perm = np.random.RandomState(0).permutation(n)
    for p in (perm, None):
        eng.set_epoch_perm(p)
        Xs = X[p] if p is not None else X
        cs = cond[p] if p is not None else cond
        for start, B in ((0, 64), (64, 1500), (2300, 57)):
            Bq, nnz = eng.gather_batch(start, B)
            torch.cuda.synchronize()
            sub = Xs[start:start + B]
            assert nnz == sub.nnz
            assert eng.indptr[: B + 1].cpu().numpy().tolist() == sub.indptr.tolist()
            assert eng.indices[:nnz].cpu().numpy().tolist() == sub.indices.tolist()
            np.testing.assert_array_equal(eng.cond[:B].cpu().numpy(), cs[start:start + B])


def test_full_ranking_k_none_matches_reference_argtopk():
    """argtopk(remove_non_missing(predict(X), X), k=None) (evaluation.py:48-52): every item of every row in descending
    order; also a k beyond the selection kernels' envelope.  Compared with the oracle's chain outside logit near-ties;
    known items come last."""
    from aaerec_b200.synth import synth_sets
    V, B = 6000, 50
    O, oracle, model, X = _steps_vs_oracle(V=V, B=B, steps=1, mean_len=9)
    Xq = synth_sets(37, V, 12, 1, 60, seed=5)
    dense = Xq.toarray()
    logits = oracle.logits(dense)
    full = model.predict_topk(Xq, None)
    assert full.shape == (37, V) and full.dtype == np.int64
    assert np.array_equal(np.sort(full, axis=1), np.tile(np.arange(V), (37, 1)))       # a permutation per row
    rows = np.arange(37)[:, None]
    n_known = dense.sum(1).astype(int)
    for r in range(37):
        assert set(full[r, V - n_known[r]:]) == set(np.nonzero(dense[r])[0])           # known items at the bottom
        head = logits[r, full[r, :V - n_known[r]]]
        assert np.all(head[:-1] >= head[1:] - 2e-6 * np.abs(head[1:]) - 1e-7)           # descending by the oracle's logits
    k = 5000                                                                            # > MAX_TOPK: same path, k columns
    top = model.predict_topk(Xq, k)
    assert np.array_equal(top, full[:, :k])
    ref = O.rank_topk(oracle.predict(dense), dense, 100)
    mism = full[:, :100] != ref
    assert mism.mean() < 0.02
    assert np.all(np.abs(logits[rows, full[:, :100]][mism] - logits[rows, ref][mism]) <= 2e-6 * np.abs(logits[rows, ref][mism]) + 1e-7)
