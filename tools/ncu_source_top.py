#!/usr/bin/env python
"""Hot SASS instructions of an .ncu-rep source page: samples, stall reasons, source line (needs -lineinfo).

    python tools/ncu_source_top.py report.ncu-rep [--top 40] [--launch 0]
"""
import csv, io, subprocess, sys
path = sys.argv[1]
top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
blocks, cur = [], []
for line in out.splitlines():
    if line.startswith('"Kernel Name"'):
        if cur: blocks.append(cur)
        cur = [line]
    elif cur:
        cur.append(line)
if cur: blocks.append(cur)
li = int(sys.argv[sys.argv.index("--launch") + 1]) if "--launch" in sys.argv else 0
b = blocks[li]
print(b[0][:160])
rows = list(csv.reader(io.StringIO("\n".join(b[1:]))))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
for n, r in enumerate(rows[1:]):
    if len(r) < len(hdr): continue
    try: s = int(r[ix["# Samples"]])
    except ValueError: continue
    data.append((s, n, r))
total = sum(d[0] for d in data) or 1
print("total samples", total)
for s, n, r in sorted(data, reverse=True)[:top]:
    st = sorted(((int(r[ix[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:3]
    print(f"{100*s/total:5.1f}% #{n:5d} {r[ix['Source']].strip()[:70]:70s} exec {r[ix['Instructions Executed']]:>9s} | " +
          " ".join(f"{c}:{v}" for v, c in st if v))
