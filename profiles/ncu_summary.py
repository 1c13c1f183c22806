#!/usr/bin/env python
"""Condenses `ncu -i X.ncu-rep --page raw --csv` into the few numbers DESIGN.md quotes per kernel.
usage: ncu -i X.ncu-rep --page raw --csv | python profiles/ncu_summary.py"""
import csv
import sys

KEEP = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "sm__cycles_elapsed.max", "sm__inst_executed.avg.per_cycle_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warp_latency_per_inst_issued.ratio", "smsp__cycles_active.avg", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active"]
rows = list(csv.reader(sys.stdin))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
names, units = rows[hdr], rows[hdr + 1]
col = {n: i for i, n in enumerate(names)}
for r in rows[hdr + 2:]:
    if len(r) < len(names):
        continue
    print("== %s  (ncu --set full --clock-control none, cold caches, one launch)" % r[col["Kernel Name"]].split("(")[0])
    for k in KEEP:
        if k in col:
            print("   %s [%s] = %s" % (k, units[col[k]], r[col[k]]))
    stalls = []
    for n, i in col.items():
        if n.startswith("smsp__average_warps_issue_stalled_") and n.endswith("_per_issue_active.ratio"):
            try:
                stalls.append((float(r[i].replace(",", "")), n[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
            except ValueError:
                pass
    stalls.sort(reverse=True)
    print("   stall cycles per issued instruction: " + ", ".join("%s %.2f" % (n, v) for v, n in stalls[:8]))
