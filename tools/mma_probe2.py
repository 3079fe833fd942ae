"""tcgen05.mma probe, round 2: cost of SS-form MMAs whose descriptors change from one MMA to the next (the convolution
kernels' tap loop) against the same descriptor repeated.  python tools/mma_probe2.py"""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import defectdetection_viaobjectdetection_b200 as paut
from defectdetection_viaobjectdetection_b200._lib import check
ctx = paut.get_context(torch.device("cuda:0"))
out = torch.zeros(128 * 16, device="cuda")
def run(mode, N, lbo, alt, reps=4000):
    check(ctx.lib.paut_debug_mma(ctx.handle, mode, N, reps, lbo, alt, C.c_void_p(out.data_ptr())), ctx.handle)
    return out[0].item()
for N in (16, 32, 64, 128, 256):
    print(f"SS N={N:3d}: same descriptor {run(4, N, 2304, 0):6.1f} | "
          + " | ".join(f"{name} mask {m}: {run(mode, N, 2304, m):6.1f}"
                       for mode, name in ((6, 'A+16B'), (7, 'A+8KB'), (8, 'B+2KB'), (9, 'A+16B,B+2KB')) for m in (1, 7)))
