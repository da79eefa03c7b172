"""Launch the decoder-output training kernel a few times on the PubMed shape (for ncu captures)."""
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
impl = int(os.environ.get("K3_IMPL", 1))
iters = int(os.environ.get("K3_ITERS", 5))
g = torch.Generator().manual_seed(0)
W = (torch.rand(V, H, generator=g) * 0.2 - 0.1).cuda()
b = torch.zeros(V).cuda()
mW, vW, mb, vb = torch.zeros_like(W), torch.zeros_like(W), torch.zeros_like(b), torch.zeros_like(b)
X = synth_sets(B, V, 16, seed=1)
ip = torch.as_tensor(X.indptr.astype(np.int32)).cuda()
ii = torch.as_tensor(X.indices.astype(np.int32)).cuda()
state = torch.zeros(48, dtype=torch.uint8).cuda()
N.call("aae_step_state_init", N.ptr(state), 1e-3, 1e-3, 0, None)
h2 = torch.relu(torch.randn(B, H, generator=g)).cuda()
dh2 = torch.zeros(B, H).cuda()
loss = torch.zeros(1, dtype=torch.float64).cuda()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
for i in range(iters):
    N.call("aae_step_tick", N.ptr(state), None)
    ev[i].record()
    N.call("aae_dec_out_train", N.ptr(h2), B, H, N.ptr(W), N.ptr(b), N.ptr(mW), N.ptr(vW), N.ptr(mb), N.ptr(vb), 0, V,
           N.ptr(ip), N.ptr(ii), float(B) * V, N.ptr(state), N.ptr(dh2), N.ptr(loss), impl, None)
ev[iters].record()
torch.cuda.synchronize()
print("impl", impl, "V", V, "B", B, "ms per launch:", [round(ev[i].elapsed_time(ev[i + 1]), 4) for i in range(iters)])
