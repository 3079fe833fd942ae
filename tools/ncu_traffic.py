#!/usr/bin/env python
"""DRAM traffic per A-scan of the fused kernels, from an `ncu --set full` capture of tools/ncu_capture.py.

    python tools/ncu_traffic.py gpurun_out/r2q_full.ncu-rep --ascans 200000 --out profiles/r02/traffic.json

Writes {profile name: {"bytes_per_ascan", "read", "write", "launches", "duration_us", "source"}} - the profile names
are the ones Ctx::launched() records and bench.py's roofline uses.  A kernel launched several times in the capture
(k_conv_tc once per layer, the attention block twice per MSC forward) is averaged over its launches: every launch
processes all the A-scans of the forward, so the figure is bytes per A-scan per LAUNCH, like roofline.achieved.
"""
import argparse
import csv
import io
import json
import os
import subprocess
import sys

# kernel function name -> Ctx::launched() name
NAMES = {
    "k_msc_encoder_tc": "msc_encoder_tc", "k_msc_attn_block": "msc_attn_block", "k_msc_ffn_head": "msc_ffn_head",
    "k_msc_attn_tc": "msc_attn_tc", "k_msc_attn_block_p": "msc_attn_block", "k_ts_encoder": "ts_encoder", "k_mscn_front": "mscn_front", "k_conv_tc": "conv_tc",
    "k_stem_flat_t": "stem_flat", "k_stem_flat": "stem_flat", "k_linear_tc": "linear_tc", "k_two_stage_final": "two_stage_final",
}


def num(s):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return 0.0


def to_bytes(v, unit):
    u = unit.strip().lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}.get(u, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--ascans", type=int, required=True, help="A-scans per forward in the capture")
    ap.add_argument("--passes", default="", help="name=count,...: for kernels whose forward was chunked in the capture, the number "
                    "of launches that together cover all A-scans once per layer (conv_tc=2: two layers), default = launches")
    ap.add_argument("--out", default="")
    ap.add_argument("--merge", action="store_true", help="keep the entries of --out for kernels that are not in this report")
    args = ap.parse_args()
    out = subprocess.run(["ncu", "-i", args.rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, launches = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    kn = idx["Kernel Name"]
    res = {}
    order = []
    for row in launches:
        fn = row[kn].split("(")[0].split("<")[0].split("::")[-1].strip()
        name = NAMES.get(fn)
        if not name:
            continue
        rd = to_bytes(num(row[idx["dram__bytes_read.sum"]]), units[idx["dram__bytes_read.sum"]])
        wr = to_bytes(num(row[idx["dram__bytes_write.sum"]]), units[idx["dram__bytes_write.sum"]])
        dur = num(row[idx["gpu__time_duration.sum"]])
        du = units[idx["gpu__time_duration.sum"]].strip().lower()
        dur_us = dur * {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3, "second": 1e6}.get(du, 1.0)
        e = res.setdefault(name, {"read": 0.0, "write": 0.0, "launches": 0, "duration_us": 0.0})
        e["read"] += rd
        e["write"] += wr
        e["launches"] += 1
        e["duration_us"] += dur_us
        order.append((fn, rd, wr, dur_us))
    passes = dict((kv.split("=")[0], int(kv.split("=")[1])) for kv in args.passes.split(",") if kv)
    for name, e in res.items():
        e["passes"] = passes.get(name, e["launches"])
        e["bytes_per_ascan"] = (e["read"] + e["write"]) / args.ascans / e["passes"]    # per launch, as bench.py's roofline
        e["source"] = f"ncu --set full, {os.path.basename(args.rep)}, {args.ascans} A-scans per forward, " \
                      f"dram__bytes_read.sum + dram__bytes_write.sum"
    for fn, rd, wr, d in order:
        print(f"{fn:22s} read {rd / 1e6:10.2f} MB  write {wr / 1e6:10.2f} MB  {d:10.1f} us  "
              f"{(rd + wr) / args.ascans:9.1f} B/A-scan")
    if args.out:
        if args.merge and os.path.exists(args.out):
            old = json.load(open(args.out))
            old.update(res)
            res = old
        with open(args.out, "w") as f:
            json.dump(res, f, indent=1, sort_keys=True)
        print("wrote", args.out)


if __name__ == "__main__":
    sys.exit(main())
