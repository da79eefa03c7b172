"""Pins oracle/aae_oracle.py against golden vectors produced by the unmodified reference
(oracle/make_golden.py) and against the reference's own doctest vectors."""
import numpy as np
import pytest

from helpers import AAE_CASES, AE_CASES, OPTION_CASES, DAE_CASES, load_case, group, oracle_replay, rel_err, GOLDEN
from oracle import aae_oracle as O


@pytest.mark.parametrize("name", AAE_CASES + AE_CASES + OPTION_CASES + DAE_CASES)
def test_partial_fit_matches_reference(name):
    g = load_case(name)
    model, losses, _, _ = oracle_replay(g)
    assert losses.shape == g["losses"].shape
    np.testing.assert_allclose(losses, g["losses"], rtol=2e-6, atol=1e-7)
    final = group(g, "final")
    assert len(final) == (18 if g["adversarial"] else 12) or not final
    for k, ref in final.items():
        assert rel_err(model.p[k].numpy(), ref) < 2e-6, k
    for k, ref in group(g, "abssum").items():
        got = np.abs(model.p[k].numpy().astype(np.float64)).sum()
        assert abs(got - ref) / ref < 2e-6, k
    cond = [g["cond"][:40]] if g["cond_dim"] else None
    pred = model.predict(g["X"][:40].toarray(), cond)
    np.testing.assert_allclose(pred, g["pred"], rtol=2e-5, atol=1e-7)


def test_initial_weights_match_reference_seed():
    g = load_case("aae_small_dropout")
    p = O.init_params(g["V"], g["H"], g["C"], seed=42)
    for k, ref in group(g, "init").items():
        np.testing.assert_array_equal(p[k].numpy(), ref)


def test_remove_non_missing_doctest():
    # evaluation.py:187-191
    Y = np.array([[0.6, 0.5, -1], [40, -20, 10]])
    X = np.array([[1, 0, 1], [0, 1, 0]])
    np.testing.assert_allclose(O.remove_non_missing(Y, X), [[0., 0.9375, 0.], [1., 0., 0.5]])


def test_argtopk_doctests():
    # evaluation.py:24-44
    X = np.arange(10).reshape(1, -1)
    r, c = O.argtopk(X, 3)
    assert r.tolist() == [[0]] and c.tolist() == [[9, 8, 7]]
    X = np.arange(20).reshape(2, 10)
    r, c = O.argtopk(X, 3)
    assert c.tolist() == [[9, 8, 7], [9, 8, 7]]
    assert X[r, c].tolist() == [[9, 8, 7], [19, 18, 17]]
    X = np.arange(6).reshape(2, 3)
    assert X[O.argtopk(X, 123123)].tolist() == [[2, 1, 0], [5, 4, 3]]


def test_ranking_golden():
    g = np.load(GOLDEN + "/ranking.npz")
    masked = O.remove_non_missing(g["Y"], g["Xk"])
    np.testing.assert_array_equal(masked, g["masked"])
    np.testing.assert_array_equal(O.argtopk(masked, 5)[1], g["top5"])
    np.testing.assert_array_equal(O.argtopk(masked, None)[1], g["full"])


@pytest.mark.parametrize("name", ["aae_small_dropout", "aae_small_cond"])
def test_rank_chain_matches_reference(name):
    g = load_case(name)
    masked = O.remove_non_missing(g["pred"], g["X"][:40].toarray())
    np.testing.assert_array_equal(masked, g["masked"])
    np.testing.assert_array_equal(O.argtopk(masked, g["k"])[1], g["topk"])
