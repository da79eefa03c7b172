"""Where does the end-to-end step (host CSR in, losses out) lose time against the device-resident step?"""
import argparse
import os
import sys
import time
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "aae-recommender_b200"))
import bench  # noqa: E402

args = argparse.Namespace(kernel="auto", no_graph=False)
ctx = bench.Ctx(args)
_, batches, V, B = bench.make_batches(os.environ.get("WL", "mpd"), 25)
eng = ctx.engine(V, B, batches)
dev = [tuple(torch.as_tensor(x, device=eng.dev) for x in (ip, ii)) for ip, ii, _ in batches]
sec = bench.train_leg(ctx, eng, dev, B, 20, 5)
print("device-resident: %.3f ms/step" % (sec / 20 * 1e3))
time.sleep(1.0)
for i in range(3):
    eng.train_step_host(batches[i][0], batches[i][1], None)
torch.cuda.synchronize()
time.sleep(1.0)
K = 20
t_call = []
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
w0 = time.perf_counter()
e0.record()
for i in range(K):
    a = time.perf_counter()
    eng.train_step_host(batches[(5 + i) % 25][0], batches[(5 + i) % 25][1], None)
    t_call.append(time.perf_counter() - a)
e1.record()
torch.cuda.synchronize()
w1 = time.perf_counter()
print("e2e: %.3f ms/step (events), wall %.3f ms/step; host time per call (us):" % (e0.elapsed_time(e1) / K, (w1 - w0) / K * 1e3),
      [int(x * 1e6) for x in t_call])
# the same host-entry graph replayed without the upload: is it the graph or the copy?
time.sleep(1.0)
tl = bench.step_timeline(eng, dev, B)
print("device-resident timeline:", {k: round(v) for k, v in tl.items()})

# variants: device-resident steps with at most two launches in flight (V1); plus the batch upload each step (V2)
def loop(mode, K=20):
    evs = [None, None]
    torch.cuda.synchronize()
    time.sleep(1.0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        if evs[i & 1] is not None:
            evs[i & 1].synchronize()
        ip, ii, _ = batches[(5 + i) % 25]
        if mode == "upload":
            eng.upload_csr(ip, ii)
        else:
            eng.set_batch_device(*dev[(5 + i) % 25])
        eng.train_step(B)
        ev = torch.cuda.Event()
        ev.record()
        evs[i & 1] = ev
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / K


for m in ("depth2", "upload", "depth2"):
    print(m, "%.3f ms/step" % loop(m))
torch.cuda.synchronize(); time.sleep(1.0)
print("train_leg again: %.3f ms/step" % (bench.train_leg(ctx, eng, dev, B, 20, 5) / 20 * 1e3))
time.sleep(1.0)
print("train_leg 60 steps: %.3f ms/step" % (bench.train_leg(ctx, eng, dev, B, 60, 5) / 60 * 1e3))
