#!/usr/bin/env python
"""profiles/r2_lk_issue.json from `ncu --set full` reports of the LK kernels: what bench.py puts into `roofline` for a kernel
whose bound is the warp-instruction issue rate, not HBM.
usage: python profiles/make_lk_issue.py key=report.ncu-rep[:kernel-regex] ...   (keys: lk_stereo_S1, lk_stereo_S64, ...)"""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out_path = os.path.join(ROOT, "profiles", "r2_lk_issue.json")
out = json.load(open(out_path)) if os.path.exists(out_path) else {}
for arg in sys.argv[1:]:
    key, rest = arg.split("=", 1)
    rep, _, rx = rest.partition(":")
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    names = rows[hdr]
    col = {n: i for i, n in enumerate(names)}
    # several launches may match (lk_pair_kernel is both the temporal and the stereo pair of calls): the one that executed the
    # most warp instructions is taken (the stereo pair: two passes over all levels)
    cand = [r for r in rows[hdr + 2:] if len(r) >= len(names) and (not rx or re.search(rx, r[col["Kernel Name"]]))]
    cand.sort(key=lambda r: -float(r[col["smsp__inst_executed.sum"]].replace(",", "")))
    for r in cand[:1]:
        f = lambda n: float(r[col[n]].replace(",", ""))
        unit = rows[hdr + 1]
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        dram = sum(f(n) * scale.get(unit[col[n]], 1) for n in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        stalls = []
        for n, i in col.items():
            if n.startswith("smsp__average_warps_issue_stalled_") and n.endswith("_per_issue_active.ratio"):
                try:
                    stalls.append((float(r[i].replace(",", "")), n[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        dur_us = f("gpu__time_duration.sum") * {"us": 1, "ms": 1e3, "ns": 1e-3}.get(unit[col["gpu__time_duration.sum"]], 1)
        out[key] = {
            "kernel": r[col["Kernel Name"]].split("(")[0], "report": os.path.basename(rep),
            "warp_instructions": f("smsp__inst_executed.sum"), "ncu_duration_us": dur_us,
            "issue_active_pct": f("smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "warps_active_pct": f("sm__warps_active.avg.pct_of_peak_sustained_active"),
            "registers_per_thread": f("launch__registers_per_thread"), "dram_bytes": dram,
            "sm_cycles": f("sm__cycles_elapsed.max"),
            "issue_rate_fraction_ncu": f("smsp__inst_executed.sum") / (f("sm__cycles_elapsed.max") * 148 * 4),
            "top_stalls": {n: round(v, 2) for v, n in stalls[:6]},
        }
        break
json.dump(out, open(out_path, "w"), indent=1)
print(json.dumps(out, indent=1))
