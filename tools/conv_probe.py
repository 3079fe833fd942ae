import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import defectdetection_viaobjectdetection_b200 as paut
from oracle import synth
kind = sys.argv[1] if len(sys.argv) > 1 else "ssd"
from tests.test_abi import MODELS
sd = synth.synth_state_dict(kind, seed=0)
m = MODELS[kind](dict(signal_length=320)); m.load_state_dict(sd); m = m.cuda().eval(); m.precision = "bf16"
x = torch.from_numpy(synth.synth_paut_sets(200, 50, 320, seed=1)).to(torch.bfloat16).cuda()
for _ in range(2): m(x)
torch.cuda.synchronize()
