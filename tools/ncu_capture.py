#!/usr/bin/env python
"""One forward of each fused path, for an `ncu --set full` capture (tools/r2q.sh runs it under ncu with a kernel-name
filter).  The set count is printed so tools/ncu_traffic.py can turn DRAM bytes per launch into bytes per A-scan.

    python tools/ncu_capture.py [--sets 4000] [--kinds msc,msc_n,two_stage,ssd]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

from defectdetection_viaobjectdetection_b200 import synthetic as synth  # noqa: E402
from defectdetection_viaobjectdetection_b200.modules import FACTORIES  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sets", type=int, default=4000)
    ap.add_argument("--kinds", default="msc,msc_n,two_stage,ssd")
    ap.add_argument("--n", type=int, default=50, help="A-scans per set")
    args = ap.parse_args()
    for kind in args.kinds.split(","):
        n = args.n
        m = FACTORIES[kind](dict(signal_length=320))
        m.load_state_dict(synth.synth_state_dict(kind, seed=0), strict=True)
        m = m.cuda().eval()
        m.precision = "bf16"
        x = torch.rand(args.sets, n, 320, device="cuda").to(torch.bfloat16)
        with torch.no_grad():
            m(x)
        torch.cuda.synchronize()
        print(f"captured {kind}: sets={args.sets} n={n} ascans={args.sets * n}", flush=True)


if __name__ == "__main__":
    main()
