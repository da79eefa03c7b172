"""Does the step / the decoder kernel time depend on the engine's age (steps done)?  For several ages: idle, then a burst of
5 + 20 back-to-back steps; idle, then the decoder kernel alone (3 + 10 launches); then isolated (synchronised) steps."""
import argparse
import os
import sys
import time
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "aae-recommender_b200"))
import bench  # noqa: E402

args = argparse.Namespace(kernel="auto", no_graph=False)
ctx = bench.Ctx(args)
_, batches, V, B = bench.make_batches("mpd", 25)
eng = ctx.engine(V, B, batches)
dev = [tuple(torch.as_tensor(x, device=eng.dev) for x in (ip, ii)) for ip, ii, _ in batches]


def burst(K=20, W=5):
    n = len(dev)
    for i in range(W):
        eng.set_batch_device(*dev[i % n]); eng.train_step(B)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        eng.set_batch_device(*dev[(W + i) % n]); eng.train_step(B)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / K


for target in [int(x) for x in os.environ.get("AGES", "0,30,60,100,160,260").split(",")]:
    while eng.steps_done < target:
        eng.set_batch_device(*dev[eng.steps_done % 25]); eng.train_step(B)
    torch.cuda.synchronize(); time.sleep(1.5)
    age0 = eng.steps_done
    b = burst()
    torch.cuda.synchronize(); time.sleep(1.5)
    pass
    r = bench.k3_roofline(ctx, eng, B, V)
    time.sleep(1.0)
    tl = bench.step_timeline(eng, dev, B)
    print("age %3d: burst %.3f ms/step | K3 alone %.3f ms | isolated step: K3 %.0f sweep %.0f span %.0f us"
          % (age0, b, r["ms"], tl["dec_out_train"], tl["w1_sweep"], tl["first_to_last_mark"]), flush=True)
