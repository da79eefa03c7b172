"""Timeline of one partial_fit inside the replayed CUDA graph: every kernel's block 0 writes %globaltimer at
its start and end (aae_trace_set); prints start/end relative to the step's first kernel, median over steps.
usage: python scripts/step_trace.py [--workload pubmed] [--steps 50]"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "aae-recommender_b200"))
import numpy as np, torch
import bench
from aaerec_b200 import _native as N
from aaerec_b200.engine import AAEEngine

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="pubmed")
ap.add_argument("--steps", type=int, default=50)
ap.add_argument("--no-graph", action="store_true")
a = ap.parse_args()
_, batches, V, B = bench.make_batches(a.workload, 16)
eng = AAEEngine(V, bench.H, bench.C, seed=1, max_batch=B, max_nnz=max(len(b[1]) for b in batches) + 8,
                use_graph=not a.no_graph)
from oracle import aae_oracle as O
eng.load_params(O.init_params(V, bench.H, bench.C, seed=42))
dev = [(torch.as_tensor(ip, device=eng.dev), torch.as_tensor(ii, device=eng.dev)) for ip, ii, _ in batches]
for i in range(5):
    eng.set_batch_device(*dev[i % 16]); eng.train_step(B)
torch.cuda.synchronize()
n = N.load().aae_trace_slots()
buf = torch.zeros(n, dtype=torch.int64, device=eng.dev)
N.call("aae_trace_set", N.ptr(buf))
rows = []
init = torch.zeros(n, dtype=torch.int64).reshape(-1, 2)
init[:, 0] = torch.iinfo(torch.int64).max
init = init.reshape(-1).to(eng.dev)
for i in range(a.steps):
    buf.copy_(init)
    eng.set_batch_device(*dev[i % 16]); eng.train_step(B)
    torch.cuda.synchronize()
    rows.append(buf.cpu().numpy().astype(np.float64).reshape(-1, 2))
N.call("aae_trace_set", None)
rows = np.stack(rows)                     # [steps, ids, 2]
valid = rows[0, :, 1] > 0
t0 = np.where(valid[None, :], rows[:, :, 0], np.inf).min(axis=1)[:, None, None]
rel = (rows - t0) * 1e-3
med = np.median(rel, axis=0)
order = [i for i in np.argsort(med[:, 0]) if valid[i]]
print("%-20s %9s %9s %9s" % ("kernel (block 0)", "start us", "end us", "dur us"))
for i in order:
    print("%-20s %9.1f %9.1f %9.1f" % (N.TRACE_NAMES[i], med[i, 0], med[i, 1], med[i, 1] - med[i, 0]))
print("step span (first start -> last end): %.1f us" % (med[valid][:, 1].max()))
