import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import defectdetection_viaobjectdetection_b200 as paut
from defectdetection_viaobjectdetection_b200 import streaming
from oracle import synth
from torch.profiler import profile, ProfilerActivity
sd = synth.synth_state_dict("msc", seed=0)
m = paut.MultiSignalClassifier(320, [128, 64, 32], 4); m.load_state_dict(sd); m = m.cuda().eval(); m.precision = "bf16"
x = torch.from_numpy(synth.synth_paut_sets(3334, 300, 320, seed=1)).to(torch.bfloat16).pin_memory()
sc = streaming.VolumeScanner(m, chunk_sets=256)
for _ in range(3): sc.scan(x)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    t = time.perf_counter(); sc.scan(x); torch.cuda.synchronize(); print("scan ms", (time.perf_counter() - t) * 1e3)
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=18, max_name_column_width=60))
ev = [e for e in prof.events() if e.device_type.name == "CUDA"] if hasattr(prof, "events") else []
