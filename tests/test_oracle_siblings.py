"""Pins the oracle's restatements of the sibling models (VAE, DecodingRecommender) against golden
vectors recorded from the unmodified reference (oracle/make_golden.py: aaerec/vae.py, aaerec/aae.py:461-584)."""
import numpy as np
import pytest

from helpers import (VAE_CASES, DECODER_CASES, load_sibling_case, group, oracle_replay_vae, oracle_replay_decoder,
                     rel_err)


@pytest.mark.parametrize("name", VAE_CASES)
def test_vae_oracle_matches_reference(name):
    g = load_sibling_case(name)
    model, losses, pred, init, _ = oracle_replay_vae(g)
    for k, ref in group(g, "init").items():
        np.testing.assert_array_equal(init[k].numpy(), ref)
    assert losses.shape == g["losses"].shape
    np.testing.assert_allclose(losses, g["losses"], rtol=3e-6, atol=1e-7)
    final = group(g, "final")
    assert len(final) == 10
    for k, ref in final.items():
        assert rel_err(model.p[k].numpy(), ref) < 3e-6, k
    np.testing.assert_allclose(pred, g["pred"], rtol=2e-5, atol=1e-7)


@pytest.mark.parametrize("name", DECODER_CASES)
def test_decoder_oracle_matches_reference(name):
    g = load_sibling_case(name)
    model, losses, pred, init = oracle_replay_decoder(g)
    for k, ref in group(g, "init").items():
        np.testing.assert_array_equal(init[k].numpy(), ref)
    assert losses.shape == g["losses"].shape
    np.testing.assert_allclose(losses, g["losses"], rtol=2e-6, atol=1e-7)
    final = group(g, "final")
    assert len(final) == 6
    for k, ref in final.items():
        assert rel_err(model.p[k].numpy(), ref) < 2e-6, k
    np.testing.assert_allclose(pred, g["pred"], rtol=2e-5, atol=1e-7)
