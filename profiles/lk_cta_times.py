#!/usr/bin/env python
"""Per-CTA start / end stamps of the 4-warp LK kernels (vf_lk.cu built with -DVF_LK_TIMING): where a launch's time goes.
usage (on a B200): python -m vins_fast_b200.build --variant lktiming -DVF_LK_TIMING -DVF_LK_TIMING_FINE
                   VF_LIB=vins_fast_b200/libvinsfast_b200_lktiming.so python profiles/lk_cta_times.py"""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vins_fast_b200 as vf                                   # noqa: E402
from vins_fast_b200.calib import EUROC_CAM0, EUROC_CAM1       # noqa: E402
from vins_fast_b200.synth import StereoSequence               # noqa: E402

seq = StereoSequence(0, 752, 480)
frames = [seq.frame(k) for k in range(40)]
b = vf.Batch(vf.make_config(752, 480, 150, 30, 1, EUROC_CAM0, EUROC_CAM1), 1)
lib = vf.load()
buf = (C.c_ulonglong * (2 * 256 * 4))()
tbuf = (C.c_ulonglong * 32)()
rows = {0: [], 1: []}
for k, (t, l, r) in enumerate(frames):
    if k == 10:
        lib.vf_debug_lk_timing(tbuf)
    n = b.track([t], [l], [r])
    if k >= 10:
        lib.vf_debug_lk_cta_times(buf)
        a = np.array(list(buf), np.int64).reshape(2, 256, 4)
        for kern in (0, 1):
            v = a[kern, :150]
            v = v[v[:, 1] > v[:, 0]]
            if len(v) == 0:
                continue
            t0 = v[:, 0].min()
            rows[kern].append({"span_us": (v[:, 1].max() - t0) / 1e3, "start_spread_us": (v[:, 0].max() - t0) / 1e3,
                               "dur_mean_us": float((v[:, 1] - v[:, 0]).mean() / 1e3), "dur_max_us": float((v[:, 1] - v[:, 0]).max() / 1e3),
                               "dur_p90_us": float(np.percentile(v[:, 1] - v[:, 0], 90) / 1e3),
                               "iters_mean": float(v[:, 3].mean()), "iters_max": int(v[:, 3].max()),
                               "us_per_iter_fit": float(np.polyfit(v[:, 3], (v[:, 1] - v[:, 0]) / 1e3, 1)[0]),
                               "us_intercept_fit": float(np.polyfit(v[:, 3], (v[:, 1] - v[:, 0]) / 1e3, 1)[1]),
                               "n": len(v), "sms_used": len(set(v[:, 2].tolist()))})
lib.vf_debug_lk_timing(tbuf)
tv = np.array(list(tbuf), np.float64)
phases = {str(i): {"mean_cycles": round(tv[i] / tv[16 + i], 1), "count": int(tv[16 + i])} for i in range(16) if tv[16 + i] > 0}
out = {}
for kern, name in ((0, "lk_temporal"), (1, "lk_stereo")):
    r = rows[kern]
    out[name] = {k: round(float(np.mean([x[k] for x in r])), 2) for k in r[0]}
out["phases"] = phases
print(json.dumps(out))
