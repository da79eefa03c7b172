"""Shared helpers for the tests: golden loading and the oracle replay of a golden case."""
import os

import numpy as np
import scipy.sparse as sp
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

AAE_CASES = ["aae_small_dropout", "aae_small_nodrop", "aae_small_cond", "aae_h100_dropout",
             "aae_survey_nodrop", "aae_survey_dropout"]
AE_CASES = ["ae_small_dropout", "ae_h100_cond"]      # the plain AutoEncoder (AAERecommender(adversarial=False))
OPTION_CASES = ["aae_opts_unnorm_scale_lrs"]         # non-default normalize_inputs / prior_scale / learning rates
DAE_CASES = ["dae_small_dropout", "dae_h100_cond"]   # DenoisingAutoEncoder (dae.py)


def load_case(name):
    g = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    g["X"] = sp.csr_matrix((np.ones(len(g["indices"]), dtype=np.float32), g["indices"], g["indptr"]),
                           shape=(int(g["n"]), int(g["V"])))
    g["dropout"] = tuple(float(x) for x in g["dropout"])
    for k in ("n", "V", "H", "C", "B", "epochs", "cond_dim", "k"):
        g[k] = int(g[k])
    g["adversarial"] = bool(int(g.get("adversarial", 1)))
    g["dae"] = bool(int(g.get("dae", 0)))
    g["noise_factor"] = float(g.get("noise_factor", 0.2))
    g["normalize_inputs"] = bool(int(g.get("normalize_inputs", 1)))
    g["prior_scale"] = float(g.get("prior_scale", 0.0)) or None
    g["gen_lr"] = float(g.get("gen_lr", 0.001))
    g["reg_lr"] = float(g.get("reg_lr", 0.001))
    return g


def group(g, prefix):
    return {k[len(prefix) + 1:]: v for k, v in g.items() if k.startswith(prefix + "/")}


def oracle_replay(g, record_rng=False):
    """Run oracle/aae_oracle.py over a golden case exactly as the reference's fit loop does
    (aae.py:768-837): seed, build, shuffle per epoch, slice batches, three phases per batch."""
    from oracle import aae_oracle as O
    V, H, C, B = g["V"], g["H"], g["C"], g["B"]
    cond = g.get("cond")
    torch.manual_seed(42)
    np.random.seed(42)
    adv = g["adversarial"]
    params = O.init_params(V, H, C, C + g["cond_dim"], seed=None, adversarial=adv)
    if adv:
        model = O.OracleAAE(params, n_code=C, gen_lr=g["gen_lr"], reg_lr=g["reg_lr"],
                            normalize_inputs=g["normalize_inputs"])
    elif g["dae"]:
        model = O.OracleDAE(params, n_code=C, noise_factor=g["noise_factor"])
    else:
        model = O.OracleAE(params, n_code=C)
    X = g["X"]
    losses, rngs, batches = [], [], []
    for _ in range(g["epochs"]):
        perm = O.fit_epoch_order(X.shape[0])
        Xs = X[perm]
        cs = cond[perm] if cond is not None else None
        for s in range(0, X.shape[0], B):
            xb = Xs[s:s + B]
            cb = [cs[s:s + B]] if cs is not None else None
            if g["dae"]:
                rng = O.draw_dae_rng(xb.shape[0], V, H, C, g["dropout"])
            else:
                rng = O.draw_step_rng(xb.shape[0], H, C, g["dropout"], prior_scale=g["prior_scale"], adversarial=adv)
            losses.append(model.partial_fit(xb.toarray(), cb, rng))
            if record_rng:
                rngs.append(rng)
                batches.append((xb, cb))
    return model, np.asarray(losses), rngs, batches


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))
