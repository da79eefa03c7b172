"""predict-only sweep (BASELINE configs[4]): reconstruction + masked top-100 over an MPD-shaped vocabulary for query
batches of 1k .. 64k sets.  Prints one JSON line per batch size (device-resident and end-to-end sets/s).
usage: python scripts/predict_sweep.py [--V 2000000] [--batches 1000,4000,16000,64000]   (torchrun for item shards)"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "aae-recommender_b200"))
import numpy as np, torch
import bench
from aaerec_b200.engine import AAEEngine
from aaerec_b200.synth import synth_sets

ap = argparse.ArgumentParser()
ap.add_argument("--V", type=int, default=2000000)
ap.add_argument("--batches", default="1000,4000,16000,64000")
ap.add_argument("--k", type=int, default=100)
a = ap.parse_args()
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
eng = AAEEngine(a.V, bench.H, bench.C, rank=rank, world=world, max_batch=128)
eng.init_uniform(42)
_, tf_peak, _ = bench.peaks()
stream = torch.cuda.current_stream()


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


for Bq in [int(x) for x in a.batches.split(",")]:
    rs = np.random.RandomState(Bq)
    lens = rs.choice([1, 5, 10, 25, 100], size=Bq)          # query lengths of eval/mpd/create_dev_set.py:16-17
    Xq = synth_sets(Bq, a.V, 25, 1, 100, seed=4321 + Bq)
    pr = bench.predict_leg(eng, Xq, a.k, 3 if Bq >= 16000 else 6, barrier, stream, tf_peak)
    if rank == 0:
        print(json.dumps({"metric": "top-100 predict sets/sec", "V": a.V, "batch": Bq, "k": a.k, "n_gpus": world,
                          "value": Bq / pr["sec"], "ms_per_batch": pr["sec"] * 1e3, "e2e": Bq / pr["sec_e2e"],
                          "algorithmic_tflops": 2.0 * Bq * a.V * bench.H / pr["sec"] / 1e12,
                          "fallbacks": eng.topk_fallbacks}))
if world > 1:
    dist.destroy_process_group()
