"""Print the key fields of bench JSON lines (one file per argument)."""
import json, sys
for path in sys.argv[1:]:
    print("==", path)
    try:
        lines = open(path).read().splitlines()
    except OSError as e:
        print("  missing:", e); continue
    for l in lines:
        l = l.strip()
        if not l.startswith("{"):
            if l: print("  ", l[:240])
            continue
        d = json.loads(l)
        print("  value %.4g %s  ms/step %.3f  launches %s  clocks %s" % (d["value"], d["unit"], d["ms_per_step"], d.get("gpu_launches"), d.get("clocks")))
        r = d.get("roofline", {})
        print("  e2e %.4g | roofline %s share %.2f achieved %s %s frac %s" % (d["e2e"]["value"], r.get("kernel"), r.get("share_of_step") or 0, r.get("achieved"), r.get("unit"), r.get("frac")))
        print("  kernels", d.get("kernels_ms_per_step"))
        print("  cpu", d.get("cpu_baseline"))
