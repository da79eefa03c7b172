"""Steady-state cost of the time-blocked W1 Adam for several group counts G (the sweep replays `pending` steps per row:
its cost only reaches steady state after G steps)."""
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
wl = os.environ.get("WL", "mpd")
_, batches, V, B = bench.make_batches(wl, 25)
for G in [int(x) for x in os.environ.get("GS", "8,16,32").split(",")]:
    os.environ["AAE_B200_W1_GROUPS"] = str(G)
    eng = ctx.engine(V, B, batches)
    dev = [tuple(torch.as_tensor(x, device=eng.dev) for x in (ip, ii)) for ip, ii, _ in batches]
    early = bench.train_leg(ctx, eng, dev, B, 20, 5) / 20 * 1e3
    for i in range(2 * G + 10):
        eng.set_batch_device(*dev[i % 25]); eng.train_step(B)
    torch.cuda.synchronize(); time.sleep(1.5)
    steady = bench.train_leg(ctx, eng, dev, B, 20, 3) / 20 * 1e3
    time.sleep(1.0)
    tl = bench.step_timeline(eng, dev, B)
    print("G", eng.w1_groups, "steps 6-25: %.3f ms/step; steady (after %d steps, 1.5 s idle): %.3f ms/step; sweep %.0f us, K3 %.0f us, span %.0f"
          % (early, 2 * G + 35, steady, tl["w1_sweep"], tl["dec_out_train"], tl["first_to_last_mark"]), flush=True)
    del eng
    torch.cuda.empty_cache()
    time.sleep(2.0)
