#!/usr/bin/env python
"""Compact summary of an .ncu-rep (raw page): duration, DRAM bytes, pipe utilisation, stall reasons.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--top-source 12]
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__inst_executed.avg.per_cycle_active",
    "sm__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
    "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_active.avg",
]


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    return hdr, units, rows[2:]


def main():
    path = sys.argv[1]
    hdr, units, launches = raw(path)
    idx = {h: i for i, h in enumerate(hdr)}
    for row in launches:
        name = row[idx.get("Kernel Name", 4)] if "Kernel Name" in idx else "?"
        print(f"== {name[:90]}")
        for k in KEYS:
            if k in idx:
                print(f"  {k:86s} {row[idx[k]]:>16s} {units[idx[k]]}")
        stalls = [(float(row[i].replace(',', '')), h) for h, i in idx.items()
                  if h.startswith("smsp__average_warp") and "issue_stalled" in h and h.endswith("_per_issue_active.ratio")
                  and row[i] not in ("", "n/a")]
        if not stalls:
            stalls = [(float(row[i].replace(',', '')), h) for h, i in idx.items()
                      if "issue_stalled" in h and h.endswith("per_warp_active.pct") and row[i] not in ("", "n/a")]
        print("  -- stall reasons (largest first)")
        for v, h in sorted(stalls, reverse=True)[:8]:
            short = h.split("issue_stalled_")[1].split("_per_")[0]
            print(f"     {short:40s} {v:10.3f}")
    if "--top-source" in sys.argv:
        n = int(sys.argv[sys.argv.index("--top-source") + 1])
        out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        if rows:
            h = rows[0]
            col = {c: i for i, c in enumerate(h)}
            samp = next((c for c in h if c.startswith("# Samples") or c == "Warp Stall Sampling (All Samples)"), None)
            src = next((c for c in h if c in ("Source", "SASS")), h[1])
            if samp:
                body = [r for r in rows[1:] if len(r) == len(h) and r[col[samp]].replace(',', '').isdigit()]
                body.sort(key=lambda r: -int(r[col[samp]].replace(',', '')))
                tot = sum(int(r[col[samp]].replace(',', '')) for r in body) or 1
                print(f"  -- hottest source lines by stall samples ({samp})")
                for r in body[:n]:
                    print(f"     {100.0 * int(r[col[samp]].replace(',', '')) / tot:5.1f}%  {r[col[src]][:110]}")


if __name__ == "__main__":
    main()
