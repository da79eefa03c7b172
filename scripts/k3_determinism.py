"""Run the decoder-output training kernel repeatedly from the same state and compare W/m/v bitwise between runs
(each element is produced by one thread with a fixed MMA order, so any difference is a race); reports which
tiles (CTA, tile iteration) differ."""
import os
import sys
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "aae-recommender_b200"))
from aaerec_b200 import _native as N  # noqa: E402
from aaerec_b200.synth import synth_sets  # noqa: E402

V = int(os.environ.get("K3_V", 200000))
B = int(os.environ.get("K3_B", 100))
H = 100
runs = int(os.environ.get("K3_RUNS", 30))
g = torch.Generator().manual_seed(0)
W0 = (torch.rand(V, H, generator=g) * 0.2 - 0.1).cuda()
b0 = (torch.rand(V, generator=g) * 0.02).cuda()
m0 = (torch.randn(V, H, generator=g) * 1e-3).cuda()
v0 = (torch.rand(V, H, generator=g) * 1e-6).cuda()
X = synth_sets(B, V, 16, seed=1)
ip = torch.as_tensor(X.indptr.astype(np.int32)).cuda()
ii = torch.as_tensor(X.indices.astype(np.int32)).cuda()
h2 = torch.relu(torch.randn(B, H, generator=g)).cuda()
sm = torch.cuda.get_device_properties(0).multi_processor_count


def run(impl):
    W, b, mW, vW = W0.clone(), b0.clone(), m0.clone(), v0.clone()
    mb, vb = torch.zeros_like(b), torch.zeros_like(b)
    state = torch.zeros(48, dtype=torch.uint8).cuda()
    N.call("aae_step_state_init", N.ptr(state), 1e-3, 1e-3, 0, None)
    N.call("aae_step_tick", N.ptr(state), None)
    dh2 = torch.zeros(B, H).cuda()
    loss = torch.zeros(1, dtype=torch.float64).cuda()
    N.call("aae_dec_out_train", N.ptr(h2), B, H, N.ptr(W), N.ptr(b), N.ptr(mW), N.ptr(vW), N.ptr(mb), N.ptr(vb), 0, V,
           N.ptr(ip), N.ptr(ii), float(B) * V, N.ptr(state), N.ptr(dh2), N.ptr(loss), impl, None)
    torch.cuda.synchronize()
    return W, mW, vW, b, dh2, loss


ref = run(1)
simt = run(0)
print("tc vs simt: W rel", ((ref[0] - simt[0]).norm() / simt[0].norm()).item(), "m rel",
      ((ref[1] - simt[1]).norm() / simt[1].norm()).item(), "dh2 rel", ((ref[4] - simt[4]).norm() / simt[4].norm()).item())
bad_runs = 0
for r in range(runs):
    out = run(1)
    diff = (out[0] != ref[0]) | (out[1] != ref[1]) | (out[2] != ref[2])
    rows = diff.any(dim=1).nonzero().flatten().cpu().numpy()
    bdiff = (out[3] != ref[3]).nonzero().flatten().cpu().numpy()
    if len(rows) or len(bdiff):
        bad_runs += 1
        tiles = np.unique(rows // 32)
        vs_simt = ((out[0] - simt[0]).abs().max().item(), (ref[0] - simt[0]).abs().max().item())
        print("run", r, "rows differing", len(rows), "tiles", [(int(t % sm), int(t // sm)) for t in tiles[:12]],
              "cols", np.unique(diff[rows[0]].nonzero().flatten().cpu().numpy())[:8] if len(rows) else None,
              "rows in tile", (rows[:8] % 32).tolist(), "bias rows", bdiff[:6].tolist(), "max|W-simt| run/ref", vs_simt)
print("runs with differences:", bad_runs, "of", runs)
