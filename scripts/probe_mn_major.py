"""Hardware probe: the B operand of the G2 / G3 GEMM forms MN-major in the SWIZZLE_128B_BASE32B layout
(aae_tc_selftest modes 6 / 7, descriptor variants in mode bits 4+)."""
import os
import sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "aae-recommender_b200"))
from aaerec_b200 import _native as N  # noqa: E402


def run(mode, A, Bm, dshape, split):
    D = torch.full(dshape, float("nan"), device="cuda")
    N.call("aae_tc_selftest", mode, N.ptr(A), N.ptr(Bm), N.ptr(D), split, None)
    torch.cuda.synchronize()
    return D.cpu().double()


g = torch.Generator().manual_seed(0)
for variant in [int(a) for a in sys.argv[1:]] or range(4):
    for split in (1, 3):
        print('  running variant', variant, 'split', split, flush=True)
        A = torch.randn(128, 32, generator=g); Bm = torch.randn(112, 32, generator=g)
        D = run(6 + 16 * variant, A.cuda(), Bm.cuda(), (128, 112), split)
        want = A.double() @ Bm.double().t()
        e6 = ((D - want).abs().max() / want.abs().max()).item()
        e6a = ((D[:, :32] - want[:, :32]).abs().max() / want.abs().max()).item()
        A = torch.randn(128, 128, generator=g); Bm = torch.randn(32, 128, generator=g)
        D = run(7 + 16 * variant, A.cuda(), Bm.cuda(), (128, 32), split)
        want = A.double() @ Bm.double().t()
        e7 = ((D - want).abs().max() / want.abs().max()).item()
        print("variant", variant, "split", split, "G2-form err %.3e (first N block %.3e)" % (e6, e6a), "G3-form err %.3e" % e7, flush=True)
