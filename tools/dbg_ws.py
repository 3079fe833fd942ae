import sys, os
sys.path.insert(0, os.getcwd())
import torch
from defectdetection_viaobjectdetection_b200 import synthetic as synth
from defectdetection_viaobjectdetection_b200.modules import FACTORIES
for kind in ("ssd", "enhanced"):
    m = FACTORIES[kind](dict(signal_length=320)); m.load_state_dict(synth.synth_state_dict(kind, seed=0), strict=True); m = m.cuda().eval(); m.precision="bf16"
    for B in (96, 2000, 20000):
        x = torch.rand(B, 50, 320, device="cuda").to(torch.bfloat16)
        try:
            out = m(x); torch.cuda.synchronize(); print(kind, B, "ok")
        except Exception as e:
            print(kind, B, "ERR", e)
