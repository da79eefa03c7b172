"""debug: run the same 3 steps repeatedly, snapshot engine state after every step, diff a good and a bad run"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "aae-recommender_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from oracle import aae_oracle as O
from aaerec_b200.aae import AdversarialAutoEncoder
from aaerec_b200.synth import synth_sets

V, B, steps, H, C = 6000, 1000, 3, 100, 50
params = O.init_params(V, H, C, seed=42)
X = synth_sets(B * steps, V, 12, seed=21)
torch.manual_seed(13)
rngs = []
for s in range(steps):
    rngs.append(torch.get_rng_state())
    O.draw_step_rng(B, H, C, (.2, .2))
NAMES = ("W1t", "W1_m1", "W1_v1", "W1_m2", "W1_v2", "Wd3", "bd3", "enc", "enc_m1", "enc_m2", "dec", "disc", "w1_last", "w1_claim",
         "h1pre", "a1", "a2", "zc", "h2", "dh2", "g_d2", "g_d1", "g_z", "g_e2", "g_h1", "h1pre2", "ga1", "ga2", "gg_z", "gg_e2", "gg_h1",
         "uniq", "n_uniq", "csc_off", "csc_row", "losses", "masks", "z_real", "indptr", "indices")

def run():
    model = AdversarialAutoEncoder(n_hidden=H, n_code=C, batch_size=B, verbose=False, rng="oracle", impl="simt", use_graph=False)
    model._build(V, C, params={k: v.clone() for k, v in params.items()})
    snaps = []
    for s in range(steps):
        torch.set_rng_state(rngs[s])
        model.partial_fit(X[s * B:(s + 1) * B])
        torch.cuda.synchronize()
        eng = model.engine
        snaps.append({n: getattr(eng, n).clone() for n in NAMES})
    return snaps

runs = [run() for _ in range(14)]
ref = runs[0]
sig = []
for r in runs:
    sig.append(float(r[-1]["W1t"].double().abs().sum()))
print("signatures:", sig)
groups = {}
for i, s in enumerate(sig):
    groups.setdefault(round(s, 3), []).append(i)
print("groups:", groups)
keys = list(groups)
if len(keys) >= 2:
    a, b = runs[groups[keys[0]][0]], runs[groups[keys[1]][0]]
    for s in range(steps):
        print("step", s)
        for n in NAMES:
            x, y = a[s][n], b[s][n]
            if x.dtype.is_floating_point:
                d = (x.double() - y.double()).abs()
                if float(d.max()) > 1e-6 * max(float(x.double().abs().max()), 1e-30):
                    nz = torch.nonzero(d.reshape(d.shape[0], -1).max(dim=1).values > 0).flatten() if d.dim() > 1 else torch.nonzero(d > 0).flatten()
                    print("   %-8s max diff %.3e  rows differing %d  first %s" % (n, float(d.max()), nz.numel(), nz[:12].tolist()))
            else:
                ne = (x != y)
                if bool(ne.any()) and n not in ("uniq", "csc_off", "csc_row"):
                    idx = torch.nonzero(ne.flatten()).flatten()
                    print("   %-8s int diffs %d first %s  a=%s b=%s" % (n, idx.numel(), idx[:8].tolist(), x.flatten()[idx[:8]].tolist(), y.flatten()[idx[:8]].tolist()))

    # detail: where do g_h1 (step 1) and the enc block (step 0) differ?
    d = (a[1]["g_h1"].double() - b[1]["g_h1"].double()).abs()
    print("g_h1 step1: |g_h1| max %.3e; per-column max diff (top 10):" % float(a[1]["g_h1"].abs().max()))
    cm = d.max(dim=0).values
    top = torch.argsort(cm, descending=True)[:10]
    print("   cols", top.tolist(), [float(cm[i]) for i in top])
    print("   rows with diff > 1e-10:", int((d.max(dim=1).values > 1e-10).sum()))
    for st_ in (0, 1):
        e = (a[st_]["enc"].double() - b[st_]["enc"].double()).abs()
        top = torch.argsort(e, descending=True)[:12]
        print("enc block step %d: top diffs at" % st_, top.tolist(), [float(e[i]) for i in top])
        for nm in ("enc_m1", "enc_m2"):
            e = (a[st_][nm].double() - b[st_][nm].double()).abs()
            top = torch.argsort(e, descending=True)[:6]
            print("   %s step %d: top diffs at" % (nm, st_), top.tolist(), [float(e[i]) for i in top], "max |x| %.3e" % float(a[st_][nm].abs().max()))
    for st_ in (0, 1):
        for nm in ("masks", "z_real", "a1", "g_e2", "h1pre2", "losses"):
            e = (a[st_][nm].double() - b[st_][nm].double()).abs()
            print("   %s step %d max diff %.3e (max |x| %.3e)" % (nm, st_, float(e.max()), float(a[st_][nm].abs().max())))
