#!/usr/bin/env python
"""Kernel start / end stamps (GPU nanosecond timer, VF_TRACE=1) of OVERLAPPED frames: where the frame period of the
asynchronous plan goes (the chain of frame k+1 beside the stereo pass of frame k).
usage (on a B200): python profiles/overlap_timeline.py [frames]      -> one JSON line (means over the steady-state frames)"""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["VF_TRACE"] = "1"
import vins_fast_b200 as vf                                   # noqa: E402
from vins_fast_b200.calib import EUROC_CAM0, EUROC_CAM1       # noqa: E402
from vins_fast_b200.synth import StereoSequence               # noqa: E402
import torch                                                  # noqa: E402

n_frames = int(sys.argv[1]) if len(sys.argv) > 1 else 60
W, H = 752, 480
seq = StereoSequence(0, W, H)
frames = np.stack([np.stack(seq.frame(k)[1:]) for k in range(n_frames)])          # [n][2][H][W]
d = torch.from_numpy(frames).cuda()
b = vf.Batch(vf.make_config(W, H, 150, 30, 1, EUROC_CAM0, EUROC_CAM1), 1)
fb = 2 * H * W
names = ("clear_frame pyrdown_left pad_left lk_temporal select_tracked response_nms bucket_scan bucket_scatter "
         "mask_and_max gftt_select pyrdown_right pad_right lk_stereo undistort_left").split()
tbuf = (C.c_ulonglong * 256)()
lib = vf.load()
lib.vf_batch_trace_get(b._h, tbuf, 256)
rows = []
# groups of 6 frames (one trace slot per frame % 6), enqueued back to back, then read
k = 0
for g in range(n_frames // 6):
    for j in range(6):
        o = d.data_ptr() + k * fb
        b.enqueue_device(k / 20.0, o, o + H * W, W, fb)
        k += 1
    b.fetch_counts()
    lib.vf_batch_trace_get(b._h, tbuf, 256)
    a = np.array(list(tbuf), np.uint64)[:192].reshape(6, 16, 2)[:, :14].astype(np.int64)
    if g >= 2:
        rows.append(a)
out = {"frames": len(rows) * 6}
per, gaps, chain = {n: [] for n in names}, [], []
for a in rows:
    for j in range(6):
        for i, n in enumerate(names):
            if a[j, i, 1] > 0:
                per[n].append((a[j, i, 1] - a[j, i, 0]) / 1e3)
        if j >= 1:   # frame j against frame j-1 of the same group
            gaps.append((a[j, 3, 0] - a[j - 1, 9, 1]) / 1e3)          # gftt_select(k-1) end -> lk_temporal(k) start
            chain.append((a[j, 3, 0] - a[j - 1, 3, 0]) / 1e3)        # period: lk_temporal start to start
out["kernel_us"] = {n: round(float(np.mean(v)), 2) for n, v in per.items() if v}
out["gftt_end_to_next_lk_temporal_start_us"] = round(float(np.mean(gaps)), 2)
out["period_us"] = round(float(np.mean(chain)), 2)
a = rows[-1]
t0 = a[1, 3, 0]
out["example_frames_1_2"] = {f"{n}@{j}": [round((a[j, i, 0] - t0) / 1e3, 1), round((a[j, i, 1] - t0) / 1e3, 1)] for j in (1, 2) for i, n in enumerate(names) if a[j, i, 1] > 0}
print(json.dumps(out))
