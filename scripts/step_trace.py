"""In-graph kernel timeline of one training step (globaltimer marks of every kernel's first/last block), cold and after
a sustained run, with clock / power samples; MPD shape by default."""
import argparse
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
import bench  # noqa: E402
from aaerec_b200 import _native as N  # noqa: E402

NAMES = ["prep", "sweep", "ae_fwd", "K3", "ae_bwd", "ae_wgrad", "rows1", "disc", "disc_wgrad", "gen", "gen_wgrad", "rows2",
         "finish", "bag_fwd", "catchup"]
wl = os.environ.get("WL", "mpd")
args = argparse.Namespace(kernel="auto", no_graph=False)
ctx = bench.Ctx(args)
_, batches, V, B = bench.make_batches(wl, 8)
eng = ctx.engine(V, B, batches)
dev = [tuple(torch.as_tensor(x, device=eng.dev) for x in (ip, ii)) for ip, ii, _ in batches]
nslots = N.load().aae_trace_slots()
buf = torch.zeros(nslots, dtype=torch.int64, device=eng.dev)


def step(i):
    eng.set_batch_device(*dev[i % len(dev)])
    eng.train_step(B)


def traced(label):
    pre = torch.tensor([2**62, 0] * (nslots // 2), dtype=torch.int64, device=eng.dev)
    rows = []
    for r in range(4):
        buf.copy_(pre)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); step(r); e1.record()
        torch.cuda.synchronize()
        t = buf.cpu().numpy().reshape(-1, 2)
        d = {NAMES[k]: (t[k, 1] - t[k, 0]) / 1e3 for k in range(len(NAMES)) if t[k, 1] > 0}
        t0 = min(t[k, 0] for k in range(len(NAMES)) if t[k, 1] > 0)
        t1 = max(t[k, 1] for k in range(len(NAMES)) if t[k, 1] > 0)
        rows.append((e0.elapsed_time(e1) * 1e3, (t1 - t0) / 1e3, d))
    ev, span, d = rows[-1]
    print(label, "step(event) %.0f us, first-to-last mark %.0f us |" % (ev, span), " ".join("%s %.0f" % (k, v) for k, v in d.items()), flush=True)


for i in range(5):
    step(i)
torch.cuda.synchronize()
N.call("aae_trace_set", N.ptr(buf))
for i in range(3):
    step(i)
time.sleep(2.0)
traced("cold     ")
samples, stop = [], False


def sampler():
    while not stop:
        try:
            o = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-i", "0"],
                               capture_output=True, text=True, timeout=5).stdout.strip().split(",")
            samples.append((int(float(o[0])), int(float(o[1]))))
        except Exception:
            pass
        time.sleep(0.15)


threading.Thread(target=sampler, daemon=True).start()
t0 = time.time()
n = 0
blocks = []
while time.time() - t0 < float(os.environ.get("SUSTAIN_S", 3.0)):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(50):
        step(n + i)
    e1.record()
    torch.cuda.synchronize()
    blocks.append(e0.elapsed_time(e1) / 50)
    n += 50
stop = True
print("sustained ms/step (blocks of 50):", [round(x, 3) for x in blocks[:3]], "...", [round(x, 3) for x in blocks[-3:]])
print("clock/power samples:", samples[::max(1, len(samples) // 12)])
traced("sustained")
