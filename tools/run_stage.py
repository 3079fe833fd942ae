"""Run one fused kernel stand-alone (paut_debug_stage) on a synthetic input: the command line for ncu captures.
    python tools/run_stage.py --stage 1 --sets 400        # two-stage fused encoder on 400 x 50 A-scans
    python tools/run_stage.py --stage 2 --sets 296        # MSC attention block (tcgen05) on 296 sets of 300 tokens"""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from defectdetection_viaobjectdetection_b200 import synthetic as synth
from defectdetection_viaobjectdetection_b200.modules import FACTORIES

ap = argparse.ArgumentParser()
ap.add_argument("--stage", type=int, default=1)
ap.add_argument("--sets", type=int, default=400)
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
kind, N, S, width = {1: ("two_stage", 50, 320, 128), 6: ("msc_n", 300, 320, 320)}.get(a.stage, ("msc", 300, 320, 64))  # stages 2-5: attention blocks
m = FACTORIES[kind](dict(signal_length=S))
m.load_state_dict(synth.synth_state_dict(kind, seed=0), strict=True)
m = m.cuda().eval()
m.precision = "bf16"
if a.stage in (1, 6):
    x = torch.from_numpy(synth.synth_paut_sets(a.sets, N, S, seed=1, defect_frac=0.01)).to(torch.bfloat16).cuda()
    native = m._native_for(x)
else:
    native = m._native_for(torch.zeros(1, 4, S, dtype=torch.bfloat16, device="cuda"))
    x = torch.randn(a.sets, N, 64, device="cuda")
for _ in range(a.reps):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = native.debug_stage(a.stage, x, width)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
print(f"stage {a.stage}: {a.sets * N} rows in {dt * 1e3:.3f} ms (host-timed, includes the launch)  checksum {out.double().sum().item():.6f}")
