"""Summarise the source page of an ncu report: instructions executed and stall samples per SASS opcode and the
hottest instructions.   python tools/ncu_hot.py report.ncu-rep [top]"""
import csv, subprocess, sys, collections, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; body = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
col = {h: i for i, h in enumerate(hdr)}
tot_inst = sum(int(r[col["Instructions Executed"]]) for r in body)
tot_samp = sum(int(r[col["# Samples"]]) for r in body)
byop = collections.Counter(); sampop = collections.Counter()
for r in body:
    op = r[col["Source"]].split()[0] if not r[col["Source"]].strip().startswith("@") else r[col["Source"]].split()[1]
    op = op.split(".")[0]
    byop[op] += int(r[col["Instructions Executed"]]); sampop[op] += int(r[col["# Samples"]])
print(f"total warp instructions {tot_inst}, samples {tot_samp}")
print("opcode: share of instructions / share of stall samples")
for op, n in byop.most_common(22):
    print(f"  {op:10s} {100*n/tot_inst:5.1f} %   {100*sampop[op]/max(tot_samp,1):5.1f} %")
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {s: sum(int(r[col[s]]) for r in body) for s in stalls}
print("stall reasons (all samples):", ", ".join(f"{k[6:]} {100*v/max(tot_samp,1):.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
print(f"hottest {top} instructions by samples:")
for r in sorted(body, key=lambda r: -int(r[col["# Samples"]]))[:top]:
    s = {k[6:]: int(r[col[k]]) for k in stalls if int(r[col[k]])}
    main = max(s, key=s.get) if s else "-"
    print(f"  {int(r[col['# Samples']]):7d}  {int(r[col['Instructions Executed']]):10d}  {main:14s} {r[col['Source']].strip()[:90]}")
