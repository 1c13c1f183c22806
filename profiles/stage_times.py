#!/usr/bin/env python
"""Per-kernel device times and algorithmic HBM rates of the parity-test workloads that are not bench lines
(BASELINE.json configs 2 and 4: 1280x720 / 400 features / 4 levels, 3840x2160 / 2000 features / 5 levels).
usage (on a B200): python profiles/stage_times.py c3|c5|c1 [frames] [streams]      -> one JSON line
(streams > 1: that many identical streams batched on the GPU, BASELINE config 4 shape)"""
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
S_ = int(sys.argv[3]) if len(sys.argv) > 3 else 1
W, H, cap, md, level = WORK[name]
cam = dict(model=1, image_width=W, image_height=H, xi=0.0,
           k1=0.0, k2=0.0, p1=0.0, p2=0.0, fx=0.55 * W, fy=0.55 * W, cx=W / 2.0, cy=H / 2.0)
seq = StereoSequence(3, W, H)
frames = [seq.frame(k) for k in range(n_frames)]
cfg = vf.make_config(W, H, cap, md, 1, cam, cam, lk_max_level=level)
b = vf.Batch(cfg, S_)
import torch                                                  # noqa: E402  (pinned host memory only)
pin = [(torch.from_numpy(l).pin_memory(), torch.from_numpy(r).pin_memory()) for _, l, r in frames]
ptr = [((C.c_void_p * S_)(*([a.data_ptr()] * S_)), (C.c_void_p * S_)(*([c.data_ptr()] * S_))) for a, c in pin]
t_warm = time.perf_counter()                                  # bring the clocks up before anything is timed (short runs
while time.perf_counter() - t_warm < 1.5:                     # otherwise measure the GPU's idle clocks)
    for k in range(n_frames):
        b.track_ptrs(frames[k][0], ptr[k][0], W, ptr[k][1], W)
b.reset()
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
# kernel spans on the GPU's own nanosecond timer (first CTA start .. last CTA end), synchronous frames
b.close()
os.environ["VF_TRACE"] = "1"
b = vf.Batch(cfg, S_)
del os.environ["VF_TRACE"]
tnames = ("clear_frame pyrdown_left pad_left lk_temporal select_tracked response_nms bucket_scan bucket_scatter "
          "mask_and_max gftt_select pyrdown_right pad_right lk_stereo undistort_left").split()
tbuf = (C.c_ulonglong * 256)()
tacc, sel_acc = [], []
for k in range(n_frames):
    b.track_ptrs(frames[k][0], ptr[k][0], W, ptr[k][1], W)
    vf.load().vf_batch_trace_get(b._h, tbuf, 256)
    if k >= 3:
        a = np.array(list(tbuf), np.uint64)[:192].reshape(6, 16, 2)[k % 6, :14]
        unused = a[:, 1] == 0                                # kernels this plan does not launch
        a = a.astype(np.int64) - int(a[~unused, 0].min())
        a[unused] = 0
        tacc.append(a)
        ph = np.array(list(tbuf), np.uint64)[224 + 8:224 + 13].astype(np.int64)   # select_tracked's phase stamps (stream 0)
        sel_acc.append(np.diff(ph))
tm = np.mean(tacc, 0) / 1e3
select_phases_us = dict(zip(["introsort", "rank", "conflict_matrix", "rounds"], [round(float(v), 1) for v in np.mean(sel_acc, 0) / 1e3]))
kernel_us = {nm: round(float(tm[i, 1] - tm[i, 0]), 1) for i, nm in enumerate(tnames)}
frame_span_us = round(float(tm[:, 1].max()), 1)
npx = W * H
S = sum(4.0 ** -l for l in range(level + 1))
alg = {"pyrdown_left": S_ * (npx + npx * (S - 1)), "response_nms": S_ * (npx + 4 * npx + npx), "mask_and_max": S_ * (5 * npx + npx)}
peak = 6548.8
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
hbm = {k: {"us": kernel_us[k], "alg_bytes": int(v), "GBps": v / (kernel_us[k] * 1e-6) / 1e9,
           "frac_of_hbm_peak": v / (kernel_us[k] * 1e-6) / 1e9 / peak} for k, v in alg.items() if kernel_us.get(k)}
print(json.dumps({"workload": name, "streams": S_, "size": [W, H], "max_cnt": cap, "lk_max_level": level, "features_last": int(n[0]),
                  "e2e_sync_us_per_frame": e2e_us, "kernel_us": kernel_us, "select_tracked_phases_us": select_phases_us, "frame_span_us": frame_span_us, "stages_us": us, "streaming": hbm, "hbm_peak_GBps": peak}))
