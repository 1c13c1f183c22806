#!/usr/bin/env python
"""Aggregates `ncu --page source --print-source sass,cuda --csv` output into per-CUDA-line stall samples.
usage: ncu -i X.ncu-rep --page source --print-source sass,cuda --csv [--kernel-name regex:K] | python ncu_source_hotspots.py [topN]"""
import csv
import sys
from collections import defaultdict

top_n = int(sys.argv[1]) if len(sys.argv) > 1 else 25
rows = list(csv.reader(sys.stdin))
cur_file, h, kernel = None, None, None
agg, src = defaultdict(lambda: defaultdict(int)), {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        kernel = r[1][:40]
        continue
    if r[0] == "Line No":
        h = r
        cols = {}
        for i, c in enumerate(h):
            cols.setdefault(c, i)
        ns, ie = h.index("# Samples"), h.index("Instructions Executed")
        continue
    if h is None or len(r) <= ns or not r[0].strip().isdigit():
        continue
    try:
        n = int(r[ns] or 0)
    except ValueError:
        continue
    key = (kernel, cur_file, int(r[0]))
    src[key] = r[1]
    agg[key]["samples"] += n
    try:
        agg[key]["inst"] += int(r[ie] or 0)
    except ValueError:
        pass
    for c, i in cols.items():
        if c.startswith("stall_"):
            try:
                agg[key][c] += int(r[i] or 0)
            except ValueError:
                pass
kernels = sorted(set(k[0] for k in agg))
for kn in kernels:
    sub = {k: v for k, v in agg.items() if k[0] == kn}
    tot = sum(v["samples"] for v in sub.values()) or 1
    print(f"== {kn}: {tot} samples, {sum(v['inst'] for v in sub.values())} warp instructions")
    for key, v in sorted(sub.items(), key=lambda kv: -kv[1]["samples"])[:top_n]:
        st = sorted(((n, c[6:]) for c, n in v.items() if c.startswith("stall_") and n and "Not" not in c), reverse=True)[:3]
        print(f"{v['samples']:6d} {v['samples'] / tot * 100:5.1f}% inst={v['inst']:8d} {key[1]}:{key[2]:<4d} {src[key].strip()[:78]} {st}")
