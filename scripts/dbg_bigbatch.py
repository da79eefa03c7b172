"""debug: per-tensor relative errors of partial_fit vs the oracle at large batches"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "aae-recommender_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from helpers import rel_err
from oracle import aae_oracle as O
from aaerec_b200.aae import AdversarialAutoEncoder
from aaerec_b200.synth import synth_sets

def run(V, B, steps, impl, dropout=(.2, .2)):
    H, C = 100, 50
    params = O.init_params(V, H, C, seed=42)
    oracle = O.OracleAAE(params, n_code=C)
    model = AdversarialAutoEncoder(n_hidden=H, n_code=C, batch_size=B, dropout=dropout, verbose=False, rng="oracle", impl=impl)
    model._build(V, C, params={k: v.clone() for k, v in params.items()})
    X = synth_sets(B * steps, V, 12, seed=21)
    torch.manual_seed(13)
    for s in range(steps):
        xb = X[s * B:(s + 1) * B]
        st = torch.get_rng_state()
        model.partial_fit(xb)
        got = model.losses()
        torch.set_rng_state(st)
        want = oracle.partial_fit(xb.toarray(), None, O.draw_step_rng(B, H, C, dropout))
        print("  step", s, "loss rel", [abs(g - w) / abs(w) for g, w in zip(got, want)])
    sd = model.state_dict()
    errs = {k: rel_err(sd[k].numpy(), v.numpy()) for k, v in oracle.p.items()}
    print("V=%d B=%d steps=%d impl=%s kernel=%s" % (V, B, steps, impl, model.engine.impl_for(B)))
    for k, e in sorted(errs.items(), key=lambda kv: -kv[1])[:6]:
        print("    %-18s %.3e" % (k, e))

for impl in ("simt", "tc"):
    for B in (100, 500, 1000):
        for steps in (1, 3):
            run(6000, B, steps, impl)
run(6000, 1000, 3, "tc", dropout=(0, 0))
