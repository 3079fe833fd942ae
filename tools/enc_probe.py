import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import defectdetection_viaobjectdetection_b200 as paut
from oracle import synth
sd = synth.synth_state_dict("msc", seed=0)
m = paut.MultiSignalClassifier(320, [128, 64, 32], 4); m.load_state_dict(sd); m = m.cuda().eval(); m.precision = "bf16"
x = torch.from_numpy(synth.synth_paut_sets(1024, 300, 320, seed=1)).to(torch.bfloat16).cuda()
for _ in range(2): m(x)
torch.cuda.synchronize()
