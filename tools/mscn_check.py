"""Full-size check of k_mscn_front (paut_debug_stage 6) against a torch fp32 reference on the GPU, in chunks.
    python tools/mscn_check.py [--sets 3334]"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from defectdetection_viaobjectdetection_b200 import synthetic as synth
from defectdetection_viaobjectdetection_b200.modules import FACTORIES

ap = argparse.ArgumentParser()
ap.add_argument("--sets", type=int, default=3334)
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
N, S = 300, 320
sd = synth.synth_state_dict("msc_n", seed=0, signal_length=S)
m = FACTORIES["msc_n"](dict(signal_length=S)); m.load_state_dict(sd, strict=True); m = m.cuda().eval(); m.precision = "bf16"
x = torch.from_numpy(synth.synth_paut_sets(a.sets, N, S, seed=1, defect_frac=0.01)).to(torch.bfloat16).cuda()
native = m._native_for(x)
w1, b1 = sd["conv1d.0.weight"].cuda().bfloat16().float(), sd["conv1d.0.bias"].cuda().float()
w2, b2 = sd["conv1d.2.weight"].cuda().half().float(), sd["conv1d.2.bias"].cuda().float()
wb, bb = sd["background_extractor.weight"].cuda().float(), sd["background_extractor.bias"].cuda().float()
def ref(xc):
    h = F.relu(F.conv1d(xc.float().unsqueeze(1), w1, b1, padding=1))
    h = F.relu(F.conv1d(h, w2, b2, padding=1))
    return (h - F.conv1d(h, wb, bb, padding=5, groups=16)).mean(1)
prev = None
for rep in range(a.reps):
    got = native.debug_stage(6, x, S)
    torch.cuda.synchronize()
    if prev is not None:
        print("rep", rep, "bitwise equal to previous run:", bool(torch.equal(got, prev)))
    prev = got.clone()
xa = x.view(-1, S)
worst, nbad, rows = 0.0, 0, []
for i in range(0, xa.shape[0], 65536):
    r = ref(xa[i:i + 65536])
    e = (got[i:i + 65536] - r).abs()
    worst = max(worst, e.max().item())
    bad = torch.nonzero(e > 2e-2)
    nbad += bad.shape[0]
    if bad.shape[0] and len(rows) < 40:
        rows += [(int(b[0]) + i, int(b[1])) for b in bad[:10]]
print(f"A-scans {xa.shape[0]}: max abs err {worst:.3e}, elements off by more than 2e-2: {nbad}; first {rows[:20]}")
print("checksum", got.double().sum().item())
