#!/usr/bin/env python
"""Per-kernel summary of an `ncu --metrics gpu__time_duration.sum --csv` launch list.
usage: python profiles/launch_summary.py launches.csv > summary.csv"""
import csv
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
t = defaultdict(list)
for r in rows:
    name = r[4].split("(")[0].replace("void ", "").replace("vf::", "")
    t[name].append(float(r[14].replace(",", "")) / 1e3)
total = sum(sum(v) for v in t.values())
print("kernel,launches,mean_us,min_us,max_us,share_of_gpu_time")
for k, v in sorted(t.items(), key=lambda kv: -sum(kv[1])):
    print("%s,%d,%.2f,%.2f,%.2f,%.3f" % (k, len(v), sum(v) / len(v), min(v), max(v), sum(v) / total))
