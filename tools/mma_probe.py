import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import defectdetection_viaobjectdetection_b200 as paut
from defectdetection_viaobjectdetection_b200._lib import check
ctx = paut.get_context(torch.device("cuda:0"))
out = torch.zeros(128 * 16, device="cuda")
for N in (16, 32, 64, 128, 256):
    for alt in (0, 1):
        if alt and N > 256: continue
        check(ctx.lib.paut_debug_mma(ctx.handle, 0, N, 2000, 0, alt, C.c_void_p(out.data_ptr())), ctx.handle)
        print(f"N={N:3d} alt_accumulators={alt}: {out[0].item():7.1f} cycles / tcgen05.mma 128xNx16")
for lbo in (16, 32, 2048):
    check(ctx.lib.paut_debug_mma(ctx.handle, 1, 16, 1, lbo, 0, C.c_void_p(out.data_ptr())), ctx.handle)
    d = out.cpu().numpy().reshape(128, 16)
    r = np.arange(128)[:, None]; e = np.arange(8)[None]
    lo = ((r * 8 + e) % 251).astype(np.float32)
    hi = (((r + lbo // 16) * 8 + e) % 251).astype(np.float32)
    ok = np.array_equal(d[:, :8], lo) and np.array_equal(d[:, 8:], hi)
    print(f"overlapped view LBO={lbo}: {'OK' if ok else 'MISMATCH'}  row0={d[0].tolist()}")

check(ctx.lib.paut_debug_mma(ctx.handle, 2, 16, 1, 0, 0, C.c_void_p(out.data_ptr())), ctx.handle)
d = out.cpu().numpy().reshape(128, 16)
exp = ((np.arange(128)[:, None] * 16 + np.arange(16)[None]) % 251).astype(np.float32)
print("TS-form MMA (A in tensor memory via tcgen05.st):", "OK" if np.array_equal(d, exp) else "MISMATCH", d[1].tolist())
for N in (16, 32, 64, 128):
    for alt in (0, 1):
        check(ctx.lib.paut_debug_mma(ctx.handle, 3, N, 2000, 0, alt, C.c_void_p(out.data_ptr())), ctx.handle)
        print(f"TS N={N:3d} alt_accumulators={alt}: {out[0].item():7.1f} cycles / tcgen05.mma 128xNx16 (A in TMEM)")
