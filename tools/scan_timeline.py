"""Per-chunk GPU timeline of the streaming scan (CUDA events): H2D start/end, kernels end."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import defectdetection_viaobjectdetection_b200 as paut
from oracle import synth
sd = synth.synth_state_dict("msc", seed=0)
m = paut.MultiSignalClassifier(320, [128, 64, 32], 4); m.load_state_dict(sd); m = m.cuda().eval(); m.precision = "bf16"
x = torch.from_numpy(synth.synth_paut_sets(3334, 300, 320, seed=1)).to(torch.bfloat16).pin_memory()
L, chunk = 4, 256
streams = [torch.cuda.Stream() for _ in range(L)]
bufs = [torch.empty((chunk, 300, 320), dtype=torch.bfloat16, device="cuda") for _ in range(L)]
def run(do_copy=True, do_compute=True, record=False):
    evs = []
    torch.cuda.synchronize()
    t0 = torch.cuda.Event(enable_timing=True); t0.record()
    tw = time.perf_counter()
    for i, first in enumerate(range(0, 3334, chunk)):
        s = streams[i % L]; n = min(chunk, 3334 - first)
        with torch.cuda.stream(s):
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            e[0].record(s)
            xd = bufs[i % L][:n]
            if do_copy: xd.copy_(x[first:first + n], non_blocking=True)
            e[1].record(s)
            if do_compute:
                native, (o, st, (B, N, S)) = m._run(xd)
                native.postprocess(st, B, N, S, 0.5, xd.device)
            e[2].record(s)
            evs.append(e)
    host = (time.perf_counter() - tw) * 1e3
    torch.cuda.synchronize()
    total = (time.perf_counter() - tw) * 1e3
    if record:
        for i, e in enumerate(evs):
            print(f"chunk {i:2d}: h2d {t0.elapsed_time(e[0]):7.2f} -> {t0.elapsed_time(e[1]):7.2f}   kernels end {t0.elapsed_time(e[2]):7.2f}")
    return host, total
for _ in range(2): run()
print("copy+compute (host issue ms, total ms):", run(record=True))
print("copy only:", run(do_compute=False))
print("compute only:", run(do_copy=False))
