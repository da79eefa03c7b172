"""debug: hunt the flaky large-batch mismatch -- repeat identical runs, report per-tensor errors vs the oracle"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "aae-recommender_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from helpers import rel_err
from oracle import aae_oracle as O
from aaerec_b200.aae import AdversarialAutoEncoder
from aaerec_b200.synth import synth_sets

V, B, steps, H, C = 6000, int(os.environ.get("BATCH", "1000")), 3, 100, 50
params = O.init_params(V, H, C, seed=42)
SEED = int(os.environ.get("SEED", "21"))
X = synth_sets(B * steps, V, 12, seed=SEED)
oracle = O.OracleAAE(params, n_code=C)
torch.manual_seed(13)
rngs = []
for s in range(steps):
    st = torch.get_rng_state()
    rngs.append(st)
    oracle.partial_fit(X[s * B:(s + 1) * B].toarray(), None, O.draw_step_rng(B, H, C, (.2, .2)))

def run(impl, graph):
    model = AdversarialAutoEncoder(n_hidden=H, n_code=C, batch_size=B, verbose=False, rng="oracle", impl=impl, use_graph=graph)
    model._build(V, C, params={k: v.clone() for k, v in params.items()})
    for s in range(steps):
        torch.set_rng_state(rngs[s])
        model.partial_fit(X[s * B:(s + 1) * B])
    sd = model.state_dict()
    eng = model.engine
    extra = {"w1_last": eng.w1_last.cpu()}
    return sd, extra

print("ENV", {k: v for k, v in os.environ.items() if k.startswith("AAE_")})
for rep in range(int(os.environ.get("REPS", "10"))):
    for impl in ("simt",):
        for graph in (True,):
            sd, extra = run(impl, graph)
            errs = {k: rel_err(sd[k].numpy(), v.numpy()) for k, v in oracle.p.items()}
            worst = max(errs, key=errs.get)
            flag = "  <<<<<< BAD" if errs[worst] > 1e-5 else ""
            print("rep %d impl %-4s graph %d worst %-18s %.3e%s" % (rep, impl, graph, worst, errs[worst], flag), flush=True)
            if flag:
                W = sd["enc.lin1.weight"].numpy().T          # [V,H]
                Wo = oracle.p["enc.lin1.weight"].numpy().T
                d = np.abs(W - Wo).max(axis=1)
                bad = np.nonzero(d > 1e-6)[0]
                print("    bad rows of W1t:", bad[:20], "n =", len(bad), "max diff", d.max())
                for r in bad[:5]:
                    cnt = [int((X[s * B:(s + 1) * B].indices == r).sum()) for s in range(steps)]
                    print("      row", r, "occurrences per step", cnt, "last", int(extra["w1_last"][r]))
                for k, e in sorted(errs.items(), key=lambda kv: -kv[1])[:5]:
                    print("      %-18s %.3e" % (k, e))
