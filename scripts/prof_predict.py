"""Per-entry-point device times of the predict / top-k path (eager, CUDA events) at a given shape."""
import os
import sys
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "aae-recommender_b200"))
from aaerec_b200 import _native as N  # noqa: E402
from aaerec_b200.engine import AAEEngine  # noqa: E402
from aaerec_b200.synth import synth_sets  # noqa: E402

V = int(os.environ.get("P_V", 2000000))
B = int(os.environ.get("P_B", 1000))
k = int(os.environ.get("P_K", 100))
iters = int(os.environ.get("P_ITERS", 3))
eng = AAEEngine(V, 100, 50, max_batch=128, impl=os.environ.get("P_IMPL", "auto"))
eng.init_uniform(42)
Xq = synth_sets(B, V, 25, 1, 100, seed=4321)
eng.upload_csr(Xq.indptr.astype(np.int32), Xq.indices.astype(np.int32))
scratch = torch.empty(B, V, dtype=torch.float32, device=eng.dev)
eng.topk(B, k, scratch=scratch)
torch.cuda.synchronize()
N.enable_timing(True)
torch.cuda.profiler.start()          # ncu --profile-from-start off: only the query loop is captured
for _ in range(iters):
    eng.topk(B, k, scratch=scratch)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
rep = N.timing_report()
N.enable_timing(False)
print("predict V=%d B=%d k=%d" % (V, B, k))
for name, (c, us) in sorted(rep.items(), key=lambda kv: -kv[1][1]):
    print("  %-24s x%d  %10.1f us each" % (name, c // iters, us))
