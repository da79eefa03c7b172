import argparse, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "aae-recommender_b200"))
import bench
args = argparse.Namespace(kernel="auto", no_graph=False)
ctx = bench.Ctx(args)
_, batches, V, B = bench.make_batches("mpd", 25)
eng = ctx.engine(V, B, batches)
dev = [tuple(torch.as_tensor(x, device=eng.dev) for x in (ip, ii)) for ip, ii, _ in batches]
for target in [1, 10, 20, 30, 35, 40, 45, 50, 55, 60, 70, 100, 200]:
    while eng.steps_done < target:
        eng.set_batch_device(*dev[eng.steps_done % 25]); eng.train_step(B)
    torch.cuda.synchronize()
    z = eng.h2[:B] @ eng.Wd3.t() + eng.bd3
    pass
    r = bench.k3_roofline(ctx, eng, B, V)
    print("age %3d: z min %.2f max %.2f mean %.2f |z|>=16: %.4f%%  rows with any: %d  h2 max %.2f  losses %s  K3 alone %.3f ms"
          % (eng.steps_done, z.min().item(), z.max().item(), z.mean().item(), (z.abs() >= 16).float().mean().item() * 100,
             int(((z.abs() >= 16).any(dim=1)).sum()), eng.h2[:B].max().item(), [round(float(x), 4) for x in eng.losses[:3].tolist()], r["ms"]), flush=True)
