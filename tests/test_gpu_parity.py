"""GPU parity tests (run on the B200 box): the CUDA path, called through the C ABI, against
(a) golden vectors produced by the unmodified reference and (b) the CPU oracle on seeded inputs."""
import numpy as np
import pytest
import torch

from helpers import AAE_CASES, AE_CASES, DAE_CASES, load_case, group, oracle_replay, rel_err

pytestmark = pytest.mark.gpu

LOSS_RTOL = 1e-4      # stated tolerance of the north star: 1e-4 relative on losses ...
WEIGHT_RTOL = 1e-4    # ... and on every weight tensor (relative Frobenius norm) after N steps
IMPLS = ["simt", "tc"]


def _make_model(g, impl, **kw):
    from aaerec_b200.aae import AdversarialAutoEncoder, AutoEncoder
    from aaerec_b200.condition import ConditionList, PrecomputedEmbeddingCondition
    conditions = None
    if g["cond_dim"]:
        conditions = ConditionList([("title", PrecomputedEmbeddingCondition(g["cond_dim"]))])
    if g["dae"]:
        from aaerec_b200.dae import DenoisingAutoEncoder
        return DenoisingAutoEncoder(n_hidden=g["H"], n_code=g["C"], batch_size=g["B"], n_epochs=g["epochs"],
                                    dropout=g["dropout"], noise_factor=g["noise_factor"], conditions=conditions,
                                    verbose=False, rng="oracle", impl=impl, **kw)
    if not g["adversarial"]:
        return AutoEncoder(n_hidden=g["H"], n_code=g["C"], batch_size=g["B"], n_epochs=g["epochs"],
                           dropout=g["dropout"], conditions=conditions, verbose=False, rng="oracle", impl=impl, **kw)
    return AdversarialAutoEncoder(n_hidden=g["H"], n_code=g["C"], batch_size=g["B"], n_epochs=g["epochs"],
                                  dropout=g["dropout"], conditions=conditions, verbose=False, rng="oracle",
                                  impl=impl, **kw)


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("name", AAE_CASES + AE_CASES + DAE_CASES)
def test_fit_matches_reference_golden(name, impl, capsys):
    """Whole fit loop (shuffle, ragged last batch, three phases, four Adam states) against the
    reference's recorded losses and final weights."""
    g = load_case(name)
    torch.manual_seed(42)
    np.random.seed(42)
    model = _make_model(g, impl)
    model.record_losses = True
    model.fit(g["X"], condition_data=[g["cond"]] if g["cond_dim"] else None)
    losses = np.asarray(model.loss_history)
    if not g["adversarial"]:       # the plain AutoEncoder logs (R, 0, 0) (aae.py:341-342)
        assert np.all(losses[:, 1:] == 0)
        losses = losses[:, :1]
    assert losses.shape == g["losses"].shape
    np.testing.assert_allclose(losses, g["losses"], rtol=LOSS_RTOL, atol=0)
    sd = model.state_dict()
    for k, ref in group(g, "final").items():
        assert rel_err(sd[k].numpy(), ref) < WEIGHT_RTOL, (k, rel_err(sd[k].numpy(), ref))
    for k, ref in group(g, "abssum").items():
        got = np.abs(sd[k].numpy().astype(np.float64)).sum()
        assert abs(got - ref) / ref < WEIGHT_RTOL, k
    # predict: dense probabilities as the reference returns them
    cond = [g["cond"][:40]] if g["cond_dim"] else None
    pred = model.predict(g["X"][:40], condition_data=cond)
    np.testing.assert_allclose(pred, g["pred"], rtol=2e-4, atol=1e-6)
    # ranking tail: identical indices outside reference score ties
    top = model.predict_topk(g["X"][:40], g["k"], condition_data=cond)
    _assert_topk_equal_outside_ties(top, g["topk"], g["masked"])


def _assert_topk_equal_outside_ties(got, ref_idx, ref_scores, tol=0.0):
    """Position-wise equality, except inside groups of equal reference scores where only the
    score sequence has to match."""
    assert got.shape == ref_idx.shape
    rows = np.arange(got.shape[0])[:, None]
    s_got = ref_scores[rows, got]
    s_ref = ref_scores[rows, ref_idx]
    mism = got != ref_idx
    # wherever indices differ the reference scores must be (nearly) tied
    assert np.all(np.abs(s_got[mism] - s_ref[mism]) <= tol + 1e-6 * np.abs(s_ref[mism])), \
        "top-k differs outside ties: %d positions" % int(mism.sum())
    assert mism.mean() < 0.05


@pytest.mark.parametrize("impl", IMPLS)
def test_twenty_steps_vs_oracle_no_dropout(impl):
    """Gate (i): dropout (0,0), 20 steps, every loss and every weight tensor within 1e-4 of the oracle."""
    from aaerec_b200.synth import synth_sets
    from oracle import aae_oracle as O
    V, H, C, B, steps = 3000, 100, 50, 100, 20
    X = synth_sets(B * steps, V, 10, seed=11)
    params = O.init_params(V, H, C, seed=42)
    oracle = O.OracleAAE(params, n_code=C)
    from aaerec_b200.aae import AdversarialAutoEncoder
    model = AdversarialAutoEncoder(n_hidden=H, n_code=C, batch_size=B, dropout=(0, 0), verbose=False, rng="oracle",
                                   impl=impl)
    model._build(V, C, params={k: v.clone() for k, v in params.items()})
    torch.manual_seed(7)
    for s in range(steps):
        xb = X[s * B:(s + 1) * B]
        st = torch.get_rng_state()
        model.partial_fit(xb)
        got = model.losses()
        torch.set_rng_state(st)
        want = oracle.partial_fit(xb.toarray(), None, O.draw_step_rng(B, H, C, (0, 0)))
        np.testing.assert_allclose(got, want, rtol=LOSS_RTOL)
    sd = model.state_dict()
    for k, v in oracle.p.items():
        assert rel_err(sd[k].numpy(), v.numpy()) < WEIGHT_RTOL, (k, rel_err(sd[k].numpy(), v.numpy()))


@pytest.mark.parametrize("impl", IMPLS)
def test_steps_vs_oracle_dropout_cond_ragged(impl):
    """Gate (ii): dropout (.2,.2) with injected oracle draws, a 300-d style condition, ragged batch sizes."""
    from aaerec_b200.synth import synth_sets, synth_condition
    from aaerec_b200.condition import ConditionList, PrecomputedEmbeddingCondition
    from aaerec_b200.aae import AdversarialAutoEncoder
    from oracle import aae_oracle as O
    V, H, C, D = 2500, 100, 50, 300
    sizes = [100, 37, 100, 1, 64, 100]
    X = synth_sets(sum(sizes), V, 12, seed=3)
    cond = synth_condition(sum(sizes), D, seed=5, scale=0.3)
    params = O.init_params(V, H, C, C + D, seed=42)
    oracle = O.OracleAAE(params, n_code=C)
    conditions = ConditionList([("title", PrecomputedEmbeddingCondition(D))])
    model = AdversarialAutoEncoder(n_hidden=H, n_code=C, batch_size=100, conditions=conditions, verbose=False,
                                   rng="oracle", impl=impl)
    model._build(V, C + D, params={k: v.clone() for k, v in params.items()})
    torch.manual_seed(9)
    s0 = 0
    for B in sizes:
        xb, cb = X[s0:s0 + B], cond[s0:s0 + B]
        s0 += B
        st = torch.get_rng_state()
        model.partial_fit(xb, condition_data=[cb])
        got = model.losses()
        torch.set_rng_state(st)
        want = oracle.partial_fit(xb.toarray(), [cb], O.draw_step_rng(B, H, C, (.2, .2)))
        np.testing.assert_allclose(got, want, rtol=LOSS_RTOL)
    sd = model.state_dict()
    for k, v in oracle.p.items():
        assert rel_err(sd[k].numpy(), v.numpy()) < WEIGHT_RTOL, (k, rel_err(sd[k].numpy(), v.numpy()))


def test_edge_cases_empty_rows_unnormalized_prior_scale():
    """Empty sets (normalize -> zero input, output = b1), normalize_inputs=False, prior_scale, odd sizes."""
    import scipy.sparse as sp
    from aaerec_b200.synth import synth_sets
    from aaerec_b200.aae import AdversarialAutoEncoder
    from oracle import aae_oracle as O
    V, H, C, B = 333, 36, 12, 17
    X = synth_sets(B, V, 4, seed=2).tolil()
    X[3, :] = 0
    X[16, :] = 0
    X = X.tocsr()
    X.eliminate_zeros()
    for normalize in (True, False):
        params = O.init_params(V, H, C, seed=1)
        oracle = O.OracleAAE(params, n_code=C, gen_lr=2e-3, reg_lr=5e-4, normalize_inputs=normalize)
        model = AdversarialAutoEncoder(n_hidden=H, n_code=C, batch_size=B, gen_lr=2e-3, reg_lr=5e-4, prior_scale=2.5,
                                       normalize_inputs=normalize, verbose=False, rng="oracle", impl="simt")
        model._build(V, C, params={k: v.clone() for k, v in params.items()})
        torch.manual_seed(3)
        for _ in range(3):
            st = torch.get_rng_state()
            model.partial_fit(X)
            got = model.losses()
            torch.set_rng_state(st)
            want = oracle.partial_fit(X.toarray(), None, O.draw_step_rng(B, H, C, (.2, .2), prior_scale=2.5))
            np.testing.assert_allclose(got, want, rtol=LOSS_RTOL)
        sd = model.state_dict()
        for k, v in oracle.p.items():
            assert rel_err(sd[k].numpy(), v.numpy()) < WEIGHT_RTOL, k


def test_graph_replay_equals_eager():
    """The CUDA-graph replay of a step must give the same weights as eager launches."""
    from aaerec_b200.synth import synth_sets
    from aaerec_b200.aae import AdversarialAutoEncoder
    from oracle import aae_oracle as O
    V, H, C, B = 1200, 100, 50, 100
    X = synth_sets(4 * B, V, 9, seed=21)
    params = O.init_params(V, H, C, seed=42)
    out = []
    for use_graph in (False, True):
        m = AdversarialAutoEncoder(n_hidden=H, n_code=C, batch_size=B, verbose=False, rng="native", seed=5,
                                   use_graph=use_graph)
        m._build(V, C, params={k: v.clone() for k, v in params.items()})
        for s in range(4):
            m.partial_fit(X[s * B:(s + 1) * B])
        out.append((m.losses(), m.state_dict()))
    np.testing.assert_allclose(out[0][0], out[1][0], rtol=1e-5)
    for k in out[0][1]:
        assert rel_err(out[1][1][k].numpy(), out[0][1][k].numpy()) < 1e-5, k


def test_native_rng_statistics():
    """In-kernel Philox dropout keeps ~(1-p) of the units and the prior sample is ~N(0, scale^2)."""
    from aaerec_b200.engine import AAEEngine
    from aaerec_b200.synth import synth_sets
    from oracle import aae_oracle as O
    V, H, C, B = 500, 100, 50, 128
    eng = AAEEngine(V, H, C, dropout=(.2, .5), prior_scale=2.0, use_graph=False, max_batch=B)
    eng.load_params(O.init_params(V, H, C, seed=3))
    X = synth_sets(B, V, 8, seed=1)
    eng.upload_csr(X.indptr.astype(np.int32), X.indices.astype(np.int32))
    # the step gathers h1pre inside its kernels: compute it separately (same weights) for the statistics
    from aaerec_b200._native import call, ptr
    call("aae_bag_fwd", ptr(eng.indptr), ptr(eng.indices), B, ptr(eng.W1t), ptr(eng.enc), H, 1, 0, V, 1,
         ptr(eng.h1pre), eng._stream())
    eng.train_step(B)
    torch.cuda.synchronize()
    h1 = eng.h1pre[:B]
    a1 = eng.a1[:B]
    kept = ((a1 != 0) | (h1 <= 0)).float().mean().item()     # units with positive pre-activation that survived
    pos = (h1 > 0)
    frac = ((a1 > 0) & pos).float().sum().item() / max(pos.float().sum().item(), 1)
    assert abs(frac - 0.8) < 0.03, frac
    zr = eng.disc_acts[:B, :C]
    assert abs(zr.mean().item()) < 0.15 and abs(zr.std().item() - 2.0) < 0.15
    # a second step draws different masks
    a1_first = a1.clone()
    eng.train_step(B)
    torch.cuda.synchronize()
    assert (eng.a1[:B] != a1_first).float().mean().item() > 0.05


def test_large_vocab_properties():
    """BASELINE-size checks (V=200k, B=100) through size-independent properties: (a) one step's losses match
    the oracle, (b) rows that are not in the batch move by exactly two zero-gradient Adam updates, (c) the sum
    over items of the dec.lin3 bias update direction is consistent with sigmoid(logit) - target."""
    from aaerec_b200.synth import synth_sets
    from aaerec_b200.aae import AdversarialAutoEncoder
    from oracle import aae_oracle as O
    V, H, C, B = 200000, 100, 50, 100
    X = synth_sets(2 * B, V, 16, seed=1)
    params = O.init_params(V, H, C, seed=42)
    oracle = O.OracleAAE(params, n_code=C)
    model = AdversarialAutoEncoder(n_hidden=H, n_code=C, batch_size=B, verbose=False, rng="oracle", impl="auto")
    model._build(V, C, params={k: v.clone() for k, v in params.items()})
    torch.manual_seed(1)
    for s in range(2):
        xb = X[s * B:(s + 1) * B]
        st = torch.get_rng_state()
        model.partial_fit(xb)
        got = model.losses()
        torch.set_rng_state(st)
        want = oracle.partial_fit(xb.toarray(), None, O.draw_step_rng(B, H, C, (.2, .2)))
        np.testing.assert_allclose(got, want, rtol=LOSS_RTOL)
    sd = model.state_dict()
    for k, v in oracle.p.items():
        assert rel_err(sd[k].numpy(), v.numpy()) < WEIGHT_RTOL, (k, rel_err(sd[k].numpy(), v.numpy()))
    # untouched rows of W1 never move when their moments are zero (zero gradient from the start)
    untouched = np.setdiff1d(np.arange(V), np.unique(X.indices))
    W1_0 = params["enc.lin1.weight"].numpy()
    np.testing.assert_array_equal(sd["enc.lin1.weight"].numpy()[:, untouched[:5000]], W1_0[:, untouched[:5000]])


def test_dae_native_corruption_drops_the_right_fraction():
    """native-RNG corruption (aae_batch_corrupt with in-kernel Philox): each entry of the batch survives with
    probability 1 - noise_factor, column order is kept, the same (seed, step) gives the same thinned batch."""
    from aaerec_b200.engine import AAEEngine
    from aaerec_b200.synth import synth_sets
    V, B = 20000, 400
    X = synth_sets(B, V, 30, seed=2)
    kept = []
    for rep in range(2):
        eng = AAEEngine(V, 100, 50, max_batch=B, adversarial=False, seed=5)
        eng.upload_csr(X.indptr.astype(np.int32), X.indices.astype(np.int32))
        eng.corrupt_batch(B, 0.3)
        torch.cuda.synchronize()
        ip = eng.indptr[: B + 1].cpu().numpy()
        ii = eng.indices[: int(ip[-1])].cpu().numpy()
        kept.append((ip.copy(), ii.copy()))
        frac = ip[-1] / X.nnz
        assert 0.66 < frac < 0.74, frac
        for r in range(B):
            row = ii[ip[r]:ip[r + 1]]
            full = X.indices[X.indptr[r]:X.indptr[r + 1]]
            assert np.all(np.diff(row) > 0) and set(row.tolist()) <= set(full.tolist())
    assert np.array_equal(kept[0][0], kept[1][0]) and np.array_equal(kept[0][1], kept[1][1])


def test_w1_sweep_cold_rows_bit_identical_to_full_replay():
    """The time-blocked sweep skips the W update (sqrt / reciprocal / fma) of COLD elements -- first moments of
    magnitude <= 2^-110, incl. never-touched rows -- because it cannot change W (w1_blocked.cu cold4).  Checked bit for
    bit against the full replay (aae_w1_catchup replays every pending step with the complete Adam arithmetic) on
    crafted rows: zero, tiny, denormal and negative-zero moments, weights below 2^-40, hot rows, and mixtures inside one
    float4."""
    import ctypes as C
    from aaerec_b200 import _native as N
    V, H, T = 4096, 100, 41
    g = torch.Generator().manual_seed(7)
    W = (torch.rand(V, H, generator=g) * 0.2 - 0.1)
    m1 = torch.randn(V, H, generator=g) * 1e-3
    v1 = torch.rand(V, H, generator=g) * 1e-5
    m2 = torch.randn(V, H, generator=g) * 1e-3
    v2 = torch.rand(V, H, generator=g) * 1e-5
    kind = torch.arange(V) % 8
    for m, v in ((m1, v1), (m2, v2)):
        m[kind == 0] = 0.0; v[kind == 0] = 0.0                        # never touched
        m[kind == 1] *= 1e-32                                          # cold (~1e-35), v alive
        m[kind == 2] = 4 * 2.0 ** -149                                 # parked on a denormal
        m[kind == 3] = -0.0; v[kind == 3] = 0.0                        # negative zero
        m[kind == 4] = 0.0                                             # zero moment, v alive
    m2[kind == 5] *= 1e-32                                             # cold in one optimizer only -> full path
    W[kind == 1, ::7] = 1e-20                                          # tiny weights among cold moments -> full path
    W[kind == 2, ::5] = 0.0
    m1[kind == 6, ::4] *= 1e-32                                        # cold and hot elements inside one float4
    # kind 7: ordinary hot rows
    state = torch.zeros(48, dtype=torch.uint8, device="cuda")
    ktab = torch.zeros(64 * 4, device="cuda")
    N.call("aae_step_state_init", N.ptr(state), 1e-3, 1e-3, C.c_uint64(0), None)
    for _ in range(T):
        N.call("aae_step_tick", N.ptr(state), None)
        N.call("aae_ktab_write", N.ptr(state), N.ptr(ktab), None)
    dev = lambda *ts: [t.clone().cuda() for t in ts]
    A = dev(W, m1, v1, m2, v2)
    Bc = dev(W, m1, v1, m2, v2)
    last_a = torch.zeros(V, dtype=torch.int32, device="cuda")
    last_b = torch.zeros(V, dtype=torch.int32, device="cuda")
    claim = torch.zeros(V, dtype=torch.int32, device="cuda")
    # A: flush sweep (flat walk with the cold path), every row from step 0 to T-1
    N.call("aae_w1_sweep_blocked", None, V, H, *[N.ptr(t) for t in A], N.ptr(last_a), N.ptr(state), N.ptr(ktab), 32, 1, 0,
           None)
    # B: the batch catch-up (full arithmetic), one "set" holding every item
    indptr = torch.tensor([0, V], dtype=torch.int32, device="cuda")
    indices = torch.arange(V, dtype=torch.int32, device="cuda")
    N.call("aae_w1_catchup", N.ptr(indptr), N.ptr(indices), 1, 0, V, N.ptr(claim), *[N.ptr(t) for t in Bc], N.ptr(last_b),
           H, N.ptr(state), N.ptr(ktab), None)
    torch.cuda.synchronize()
    assert torch.equal(last_a, last_b) and int(last_a.min()) == T - 1
    for name, a, b in zip(("W", "m1", "v1", "m2", "v2"), A, Bc):
        assert torch.equal(a.view(torch.int32), b.view(torch.int32)), name
    # the replay did something: hot rows moved, never-touched rows did not
    assert not torch.equal(A[0][kind == 7].cpu(), W[kind == 7])
    assert torch.equal(A[0][kind == 0].cpu(), W[kind == 0])
