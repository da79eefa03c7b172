"""Decoder-output training kernel launched back to back for several seconds: per-launch time at the start and at the
end, with the SM clock and power sampled by nvidia-smi (does the power cap slow it down?)."""
import os
import subprocess
import sys
import threading
import time
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "aae-recommender_b200"))
from aaerec_b200 import _native as N  # noqa: E402
from aaerec_b200.synth import synth_sets  # noqa: E402

V = int(os.environ.get("K3_V", 2000000))
B, H = 100, 100
iters = int(os.environ.get("K3_ITERS", 2500))
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
samples = []
stop = False


def sampler():
    while not stop:
        try:
            o = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,temperature.gpu", "--format=csv,noheader,nounits",
                                "-i", "0"], capture_output=True, text=True, timeout=5).stdout.strip().split(",")
            samples.append((time.time(), float(o[0]), float(o[1]), float(o[2])))
        except Exception:
            pass
        time.sleep(0.1)


th = threading.Thread(target=sampler, daemon=True)
th.start()
time.sleep(0.5)
iters -= iters % 50
ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters // 50 + 1)]
t0 = time.time()
for i in range(iters):
    if i % 50 == 0:
        ev[i // 50].record()
    N.call("aae_dec_out_train", N.ptr(h2), B, H, N.ptr(W), N.ptr(b), N.ptr(mW), N.ptr(vW), N.ptr(mb), N.ptr(vb), 0, V,
           N.ptr(ip), N.ptr(ii), float(B) * V, N.ptr(state), N.ptr(dh2), N.ptr(loss), 1, None)
ev[-1].record()
torch.cuda.synchronize()
t1 = time.time()
stop = True
ms = [ev[i].elapsed_time(ev[i + 1]) / 50 for i in range(len(ev) - 1)]
print("ms per launch (blocks of 50):", [round(x, 4) for x in ms[:4]], "...", [round(x, 4) for x in ms[-4:]], "wall", round(t1 - t0, 2))
busy = [s for s in samples if t0 <= s[0] <= t1]
if busy:
    print("during: sm_mhz", [int(s[1]) for s in busy[::max(1, len(busy) // 10)]], "power_w", [int(s[2]) for s in busy[::max(1, len(busy) // 10)]],
          "temp", [int(s[3]) for s in busy[::max(1, len(busy) // 5)]])
