"""Where does VolumeScanner.scan spend its time?  (host launch vs sync vs D2H), for a few chunk sizes."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import defectdetection_viaobjectdetection_b200 as paut
from defectdetection_viaobjectdetection_b200 import streaming
from oracle import synth

sd = synth.synth_state_dict("msc", seed=0)
m = paut.MultiSignalClassifier(320, [128, 64, 32], 4); m.load_state_dict(sd); m = m.cuda().eval(); m.precision = "bf16"
n_sets = 3334
x = torch.from_numpy(synth.synth_paut_sets(n_sets, 300, 320, seed=1)).to(torch.bfloat16).pin_memory()
xd = x.cuda()
for _ in range(3): m.predict_records(xd[:256])
torch.cuda.synchronize()
t = time.perf_counter(); r = m.predict_records(xd); torch.cuda.synchronize(); print("resident whole", time.perf_counter() - t, len(r))
for chunk in (128, 256, 512, 1024, 3334):
    sc = streaming.VolumeScanner(m, chunk_sets=chunk)
    sc.scan(x)
    acc = {"harvest": 0.0, "launch": 0.0}
    orig_h = sc._harvest
    def timed_h(lane, out, _o=orig_h):
        t0 = time.perf_counter(); _o(lane, out); acc["harvest"] += time.perf_counter() - t0
    sc._harvest = timed_h
    torch.cuda.synchronize(); t = time.perf_counter(); r = sc.scan(x); torch.cuda.synchronize(); dt = time.perf_counter() - t
    print({k: round(v * 1e3, 2) for k, v in sc.stats.items()})
    print(f"chunk {chunk}: total {dt*1e3:.1f} ms  harvest {acc['harvest']*1e3:.1f} ms  rate {n_sets*300/dt/1e6:.1f} M/s  records {len(r)}")
# per-chunk device time without host copies
for chunk in (256, 1024):
    xs = xd[:chunk].contiguous()
    for _ in range(3): m.predict_records(xs)
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(10): native, (o, st, (B, N, S)) = m._run(xs); native.postprocess(st, B, N, S, 0.5, xs.device)
    torch.cuda.synchronize(); print(f"device-only chunk {chunk}: {(time.perf_counter()-t)/10*1e3:.2f} ms")
    t = time.perf_counter()
    for _ in range(10): native, (o, st, (B, N, S)) = m._run(xs); native.postprocess(st, B, N, S, 0.5, xs.device)
    print(f"host launch cost chunk {chunk}: {(time.perf_counter()-t)/10*1e3:.2f} ms"); torch.cuda.synchronize()
