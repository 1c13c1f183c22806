#!/usr/bin/env python
"""Per-kernel device times and algorithmic HBM rates of the parity-test workloads that are not bench lines
(BASELINE.json configs 2 and 4: 1280x720 / 400 features / 4 levels, 3840x2160 / 2000 features / 5 levels).
usage (on a B200): python profiles/stage_times.py c3|c5|c1 [frames]      -> one JSON line"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vins_fast_b200 as vf                                   # noqa: E402
from vins_fast_b200.synth import StereoSequence               # noqa: E402

WORK = {"c1": (752, 480, 150, 30, 3), "c3": (1280, 720, 400, 30, 4), "c5": (3840, 2160, 2000, 30, 5)}
name = sys.argv[1] if len(sys.argv) > 1 else "c5"
n_frames = int(sys.argv[2]) if len(sys.argv) > 2 else 10
W, H, cap, md, level = WORK[name]
cam = dict(model=1, image_width=W, image_height=H, xi=0.0,
           k1=0.0, k2=0.0, p1=0.0, p2=0.0, fx=0.55 * W, fy=0.55 * W, cx=W / 2.0, cy=H / 2.0)
seq = StereoSequence(3, W, H)
frames = [seq.frame(k) for k in range(n_frames)]
cfg = vf.make_config(W, H, cap, md, 1, cam, cam, lk_max_level=level)
b = vf.Batch(cfg, 1)
import torch                                                  # noqa: E402  (pinned host memory only)
pin = [(torch.from_numpy(l).pin_memory(), torch.from_numpy(r).pin_memory()) for _, l, r in frames]
ptr = [((C.c_void_p * 1)(a.data_ptr()), (C.c_void_p * 1)(c.data_ptr())) for a, c in pin]
for k in range(min(7, n_frames - 1)):
    b.track_ptrs(frames[k][0], ptr[k][0], W, ptr[k][1], W)
t0 = time.perf_counter()
for k in range(min(7, n_frames - 1), n_frames):
    n = b.track_ptrs(frames[k][0], ptr[k][0], W, ptr[k][1], W)
e2e_us = (time.perf_counter() - t0) / (n_frames - min(7, n_frames - 1)) * 1e6
b.reset()
b.set_profiling(True)
acc, cnt = {}, 0
for k in range(n_frames):
    b.track_ptrs(frames[k][0], ptr[k][0], W, ptr[k][1], W)
    if k >= 3:
        for s, ms in b.stage_times_ms().items():
            acc[s] = acc.get(s, 0.0) + ms * 1e3
        cnt += 1
us = {k: v / cnt for k, v in acc.items()}
npx = W * H
S = sum(4.0 ** -l for l in range(level + 1))
alg = {"pyrdown_left": npx + npx * (S - 1), "response_nms": npx + 4 * npx + npx, "mask_and_max": 5 * npx + npx}
peak = 6548.8
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
hbm = {k: {"us": us.get(k), "alg_bytes": int(v), "GBps": v / (us[k] * 1e-6) / 1e9, "frac_of_hbm_peak": v / (us[k] * 1e-6) / 1e9 / peak}
       for k, v in alg.items() if us.get(k)}
print(json.dumps({"workload": name, "size": [W, H], "max_cnt": cap, "lk_max_level": level, "features_last": int(n[0]),
                  "e2e_sync_us_per_frame": e2e_us, "stages_us": us, "streaming": hbm, "hbm_peak_GBps": peak}))
