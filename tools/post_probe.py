"""Where does the time attributed to post_count go?  Times forward and postprocess separately with CUDA events
(no per-kernel profiling), for models whose bench line showed a multi-ms post_count."""
import sys, os, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import synth
from tests.test_abi import MODELS

dev = torch.device("cuda:0")
for kind in sys.argv[1:] or ["msc_legacy", "hybrid", "complex"]:
    sd = synth.synth_state_dict(kind, seed=0)
    m = MODELS[kind](dict(signal_length=320)); m.load_state_dict(sd); m = m.to(dev).eval(); m.precision = "bf16"
    x = torch.from_numpy(synth.synth_paut_sets(3334, 300, 320, seed=42)).to(torch.bfloat16).to(dev)
    for _ in range(3):
        native, (outs, struct, (B, N, S)) = m._run(x)
        native.postprocess(struct, B, N, S, 0.5, dev)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    t0 = time.perf_counter()
    ev[0].record()
    for _ in range(5):
        native, (outs, struct, (B, N, S)) = m._run(x)
    ev[1].record()
    t1 = time.perf_counter()
    for _ in range(5):
        det, count = native.postprocess(struct, B, N, S, 0.5, dev)
    ev[2].record()
    t2 = time.perf_counter()
    torch.cuda.synchronize()
    print(kind, "forward gpu ms/step %.3f (host %.3f)  postprocess gpu ms/step %.3f (host %.3f) kept %d" % (
        ev[0].elapsed_time(ev[1]) / 5, (t1 - t0) * 200, ev[1].elapsed_time(ev[2]) / 5, (t2 - t1) * 200, int(count.item())))
