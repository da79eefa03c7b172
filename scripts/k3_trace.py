"""Per-tile timeline of the decoder-output training kernel (variant built with -DK3X_TRACE): clock64() of CTA 0 at
the synchronisation points of tile iterations 8..39, printed as cycles relative to the MMA warp's wake-up."""
import ctypes
import os
import runpy
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ.setdefault("K3_ITERS", "4")
runpy.run_path(os.path.join(ROOT, "scripts", "prof_k3.py"), run_name="__main__")
sys.path.insert(0, os.path.join(ROOT, "aae-recommender_b200"))
from aaerec_b200 import _native as N  # noqa: E402

lib = N.load()
out = np.zeros(32 * 16, dtype=np.int64)
f = lib.aae_k3x_trace_read
f.argtypes = [ctypes.c_void_p]
f.restype = ctypes.c_int
rc = f(out.ctypes.data)
t = out.reshape(32, 16)
names = ["mma_wakeA", "mma_g1_issued", "e1_top", "e1_g1done", "e1_mathdone", "e1_g23done", "e1_arrivedB", "e2_done",
         "e2_stagewait", "e2_stageok", "e2_dwready", "-", "-", "-", "mma_wakeB", "mma_g23_issued"]
print("rc", rc, "columns:", names)
base = t[:, 0]
for r in range(1, 31):
    print(r + 8, "period", int(t[r, 0] - t[r - 1, 0]), [int(t[r, k] - base[r]) for k in range(1, 16)])
per = np.diff(t[1:31, 0])
print("mean period", per.mean(), "median", np.median(per))
rel = (t[1:31, 1:16] - t[1:31, 0:1])
print("median rel:", dict(zip(names[1:], np.median(rel, axis=0).astype(int).tolist())))

if int(os.environ.get("K3_V", 200000)) <= 32 * 148 * 30:
    pn = ["entry", "zero_filled", "hb_filled", "ht_tmem", "events", "loop_start", "loop_end", "dh2_flushed", "exit"]
    row = t[31, :9]
    print("prologue/epilogue cycles since entry:", dict(zip(pn, (row - row[0]).astype(int).tolist())))
