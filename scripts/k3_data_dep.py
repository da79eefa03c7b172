"""Is the decoder-output training kernel's time data dependent?  Same launch, cold GPU, three inputs:
uniform synthetic sets / the bench's MPD-shaped batch / the MPD batch with a dense h2."""
import os
import sys
import time
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "aae-recommender_b200"))
from aaerec_b200 import _native as N  # noqa: E402
from aaerec_b200.synth import synth_sets  # noqa: E402
import bench  # noqa: E402

V, B, H = 2000000, 100, 100
g = torch.Generator().manual_seed(0)
W = (torch.rand(V, H, generator=g) * 0.2 - 0.1).cuda()
b = torch.zeros(V).cuda()
mW, vW, mb, vb = torch.zeros_like(W), torch.zeros_like(W), torch.zeros_like(b), torch.zeros_like(b)
state = torch.zeros(48, dtype=torch.uint8).cuda()
N.call("aae_step_state_init", N.ptr(state), 1e-3, 1e-3, 0, None)
dh2 = torch.zeros(B, H).cuda()
loss = torch.zeros(1, dtype=torch.float64).cuda()
gw = torch.zeros(V * H + V).cuda()


def timed(name, ip, ii, h2, entry, iters=8):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    time.sleep(2.0)
    for i in range(iters):
        N.call("aae_step_tick", N.ptr(state), None)
        ev[i].record()
        if entry == "ws":
            N.call("aae_dec_out_train_ws", N.ptr(h2), B, H, N.ptr(W), N.ptr(b), N.ptr(mW), N.ptr(vW), N.ptr(mb), N.ptr(vb), 0, V,
                   N.ptr(ip), N.ptr(ii), float(B) * V, N.ptr(state), N.ptr(dh2), N.ptr(loss), 1, N.ptr(gw), int(gw.numel()), None)
        else:
            N.call("aae_dec_out_train", N.ptr(h2), B, H, N.ptr(W), N.ptr(b), N.ptr(mW), N.ptr(vW), N.ptr(mb), N.ptr(vb), 0, V,
                   N.ptr(ip), N.ptr(ii), float(B) * V, N.ptr(state), N.ptr(dh2), N.ptr(loss), 1, None)
    ev[iters].record()
    torch.cuda.synchronize()
    print("%-34s" % name, [round(ev[i].elapsed_time(ev[i + 1]), 4) for i in range(iters)], flush=True)


X = synth_sets(B, V, 16, seed=1)
ipu = torch.as_tensor(X.indptr.astype(np.int32)).cuda()
iiu = torch.as_tensor(X.indices.astype(np.int32)).cuda()
h2r = torch.relu(torch.randn(B, H, generator=g)).cuda()
_, batches, _, _ = bench.make_batches("mpd", 2)
ipm = torch.as_tensor(batches[0][0]).cuda()
iim = torch.as_tensor(batches[0][1]).cuda()
print("mpd batch: nnz", int(iim.numel()), "items < 4736:", int((iim < 4736).sum()), "max row", int((ipm[1:] - ipm[:-1]).max()))
timed("uniform sets, relu(randn) h2", ipu, iiu, h2r, "plain")
timed("mpd sets, relu(randn) h2", ipm, iim, h2r, "plain")
timed("mpd sets, ws entry", ipm, iim, h2r, "ws")
timed("uniform sets again", ipu, iiu, h2r, "plain")
