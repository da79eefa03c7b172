"""GPU parity of the sibling models that share the decoder output layer: the B200 VAE
(aaerec/vae.py:47-266) and DecodingRecommender (aaerec/aae.py:461-584), through the C ABI, against (a) golden vectors
recorded from the unmodified reference and (b) the CPU oracle at a PubMed-like layer shape."""
import numpy as np
import pytest
import torch

from helpers import (VAE_CASES, DECODER_CASES, load_sibling_case, group, rel_err, linear_init)
from test_gpu_parity import _assert_topk_equal_outside_ties

pytestmark = pytest.mark.gpu

LOSS_RTOL = 1e-4      # the north star's tolerance: 1e-4 relative on losses and on every weight tensor
WEIGHT_RTOL = 1e-4


def _row_condition(dim):
    from aaerec_b200.condition import ConditionList, PrecomputedEmbeddingCondition
    return ConditionList([("title", PrecomputedEmbeddingCondition(dim))])


@pytest.mark.parametrize("impl", ["simt", "tc"])
@pytest.mark.parametrize("name", VAE_CASES)
def test_vae_fit_matches_reference_golden(name, impl):
    from aaerec_b200.vae import VAE
    g = load_sibling_case(name)
    torch.manual_seed(42)
    np.random.seed(42)
    conditions = _row_condition(g["cond_dim"]) if g["cond_dim"] else None
    model = VAE(g["V"], g["V"], n_hidden=g["H"], n_code=g["C"], lr=g["lr"], batch_size=g["B"], n_epochs=g["epochs"],
                conditions=conditions, verbose=False, rng="oracle", impl=impl)
    sd0 = model.state_dict()
    for k, ref in group(g, "init").items():
        np.testing.assert_array_equal(sd0[k].numpy(), ref)
    model.record_losses = True
    cond = [g["cond"]] if g["cond_dim"] else None
    model.fit(g["X"], condition_data=cond)
    losses = np.asarray(model.loss_history)[:, 0]
    assert losses.shape == g["losses"].shape
    np.testing.assert_allclose(losses, g["losses"], rtol=LOSS_RTOL, atol=0)
    sd = model.state_dict()
    final = group(g, "final")
    assert len(final) == 10
    for k, ref in final.items():
        assert rel_err(sd[k].numpy(), ref) < WEIGHT_RTOL, (k, rel_err(sd[k].numpy(), ref))
    # predict samples in eval mode too (vae.py:252-256): replay the reference's draws
    torch.manual_seed(g["predict_seed"])
    pred = model.predict(g["X"][:40], condition_data=[g["cond"][:40]] if g["cond_dim"] else None)
    np.testing.assert_allclose(pred, g["pred"], rtol=2e-4, atol=1e-6)
    # the fused ranking tail runs on native noise: same shape, valid unknown items only
    top = model.predict_topk(g["X"][:40], g["k"], condition_data=[g["cond"][:40]] if g["cond_dim"] else None)
    assert top.shape == (40, g["k"])
    known = g["X"][:40].toarray() > 0
    assert not known[np.arange(40)[:, None], top].any()


@pytest.mark.parametrize("impl", ["simt", "tc"])
@pytest.mark.parametrize("name", DECODER_CASES)
def test_decoder_fit_matches_reference_golden(name, impl):
    from aaerec_b200.decoding import DecodingRecommender
    g = load_sibling_case(name)
    torch.manual_seed(42)
    np.random.seed(42)
    rec = DecodingRecommender(_row_condition(g["D"]), n_epochs=g["epochs"], batch_size=g["B"], n_hidden=g["H"],
                              lr=g["lr"], verbose=False, dropout=g["dropout"], rng="oracle", impl=impl)
    rec._make_model()
    rec.model.record_losses = True
    rec.model.fit(g["X"], condition_data=[g["cond"]])
    losses = np.asarray(rec.model.loss_history)
    assert np.all(losses[:, 1:] == 0)
    np.testing.assert_allclose(losses[:, 0], g["losses"], rtol=LOSS_RTOL, atol=0)
    sd = rec.model.engine.state_dict()
    final = group(g, "final")
    assert len(final) == 6
    for k, ref in final.items():
        assert rel_err(sd[k].numpy(), ref) < WEIGHT_RTOL, (k, rel_err(sd[k].numpy(), ref))
    assert rec.mlp.lin3.weight.shape == (g["V"], g["H"])

    class QueryBags(object):
        def size(self, dim):
            return 40

        def get_attributes(self, keys):
            return [g["cond"][:40]]

        def tocsr(self):
            return g["X"][:40]
    pred = rec.predict(QueryBags())
    np.testing.assert_allclose(pred, g["pred"], rtol=2e-4, atol=1e-6)
    top = rec.predict_topk(QueryBags(), g["k"])
    from oracle import aae_oracle as O
    masked = O.remove_non_missing(g["pred"], g["X"][:40].toarray())
    _assert_topk_equal_outside_ties(top, g["topk"], masked)


def test_decoder_fit_is_the_public_entry(capsys):
    """``fit(condition_data, Y)`` / ``partial_fit(condition_data, y)`` as the reference spells them (aae.py:490, 522)."""
    from aaerec_b200.decoding import DecodingRecommender
    g = load_sibling_case("decoder_small_dropout")
    torch.manual_seed(42)
    np.random.seed(42)
    rec = DecodingRecommender(_row_condition(g["D"]), n_epochs=1, batch_size=g["B"], n_hidden=g["H"], lr=g["lr"],
                              verbose=True, dropout=g["dropout"], rng="oracle")
    rec.fit([g["cond"]], g["X"])
    assert "Loss: 0.69" in capsys.readouterr().out                     # aae.py:517-518
    rec.partial_fit([g["cond"][:20]], torch.as_tensor(g["X"][:20].toarray()))
    assert rec.model.engine.steps_done == 4
    assert "MLP-2 Decoder with %d hidden units" % g["H"] in str(rec)


@pytest.mark.parametrize("impl", ["simt", "tc"])
def test_vae_steps_vs_oracle_pubmed_like(impl):
    """H=100, C=50, B=100 on a 30k-item vocabulary, 300-d row condition: 6 partial_fit steps against OracleVAE."""
    from aaerec_b200.synth import synth_sets
    from aaerec_b200.vae import VAE
    from oracle import aae_oracle as O
    V, H, C, B, D, steps = 30000, 100, 50, 100, 300, 6
    X = synth_sets(B * steps, V, 12, seed=21)
    cond = (np.random.RandomState(5).randn(B * steps, D) * 0.1).astype(np.float32)
    torch.manual_seed(42)
    params = linear_init([("fc1", V, H), ("fc21", H, C), ("fc22", H, C), ("fc3", C + D, H), ("fc4", H, V)])
    oracle = O.OracleVAE({k: v.clone() for k, v in params.items()}, n_code=C)
    model = VAE(V, V, n_hidden=H, n_code=C, batch_size=B, conditions=_row_condition(D), verbose=False, rng="oracle",
                impl=impl, params={k: v.clone() for k, v in params.items()})
    torch.manual_seed(9)
    for s in range(steps):
        xb, cb = X[s * B:(s + 1) * B], cond[s * B:(s + 1) * B]
        st = torch.get_rng_state()
        model.partial_fit(xb, condition_data=[cb])
        got = model.losses()[0]
        torch.set_rng_state(st)
        want = oracle.partial_fit(xb.toarray(), [cb], torch.randn((B, C), dtype=torch.float32))
        assert abs(got - want) / abs(want) < LOSS_RTOL, (s, got, want)
    sd = model.state_dict()
    for k, ref in oracle.p.items():
        assert rel_err(sd[k].numpy(), ref.numpy()) < WEIGHT_RTOL, (k, rel_err(sd[k].numpy(), ref.numpy()))


@pytest.mark.parametrize("B", [100, 333])
def test_decoder_steps_vs_oracle_pubmed_like(B):
    """H=100 on a 30k-item vocabulary from a 300-d condition, batch 100 (one row chunk of K3) and 333 (chunked)."""
    from aaerec_b200.synth import synth_sets
    from aaerec_b200.decoding import _DecoderNet
    from oracle import aae_oracle as O
    V, H, D, steps = 30000, 100, 300, 5
    Y = synth_sets(B * steps, V, 12, seed=22)
    cond = (np.random.RandomState(6).randn(B * steps, D) * 0.1).astype(np.float32)
    torch.manual_seed(42)
    params = linear_init([("lin1", D, H), ("lin2", H, H), ("lin3", H, V)])
    oracle = O.OracleDecoder({k: v.clone() for k, v in params.items()})
    model = _DecoderNet(n_hidden=H, batch_size=B, conditions=_row_condition(D), verbose=False, rng="oracle")
    model._build(V, D, params={k: v.clone() for k, v in params.items()})
    torch.manual_seed(9)
    for s in range(steps):
        yb, cb = Y[s * B:(s + 1) * B], cond[s * B:(s + 1) * B]
        st = torch.get_rng_state()
        model.partial_fit(yb, condition_data=[cb])
        got = model.losses()[0]
        torch.set_rng_state(st)
        masks = (O.draw_masks((B, H), .2, 1)[0], O.draw_masks((B, H), .2, 1)[0])
        want = oracle.partial_fit([cb], yb.toarray(), {"ae_dec": masks})
        assert abs(got - want) / abs(want) < LOSS_RTOL, (s, got, want)
    sd = model.engine.state_dict()
    for k, ref in oracle.p.items():
        assert rel_err(sd[k].numpy(), ref.numpy()) < WEIGHT_RTOL, (k, rel_err(sd[k].numpy(), ref.numpy()))
