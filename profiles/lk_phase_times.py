#!/usr/bin/env python
"""Cycle accounting of the warp-level LK passes (vf_lk_warp.cu built with -DVF_LKW_TIMING).
usage (on a B200):  python -m vins_fast_b200.build --timing && VF_LIB=vins_fast_b200/libvinsfast_b200_timing.so python profiles/lk_phase_times.py [streams] [lk_mode]"""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vins_fast_b200 as vf                                   # noqa: E402
from vins_fast_b200 import ops                                # noqa: E402
from vins_fast_b200.calib import EUROC_CAM0, EUROC_CAM1       # noqa: E402
from vins_fast_b200.synth import StereoSequence               # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 1
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 0
ops.set_lk_kernel(mode)
seq = StereoSequence(0, 752, 480)
frames = [seq.frame(k) for k in range(40)]
b = vf.Batch(vf.make_config(752, 480, 150, 30, 1, EUROC_CAM0, EUROC_CAM1), S)
lib = vf.load()
buf = (C.c_ulonglong * 32)()
names = ["prologue", "pro:tileI", "pro:+patches", "level", "level:setup", "level:loop", "J staged on demand", "iterations", "fast iterations",
         "coop:prologues", "coop:pass"]
for k, (t, l, r) in enumerate(frames):
    if k == 10:
        lib.vf_debug_lkw_timing(buf)
    b.track([t] * S, [l] * S, [r] * S)
lib.vf_debug_lkw_timing(buf)
v = np.array(list(buf), np.float64)
out = {"streams": S, "lk_mode": mode, "frames": 30}
for i, n in enumerate(names):
    if i in (7, 8):
        out[n] = v[i]
    elif v[16 + i] > 0:
        out[n] = {"mean_cycles": round(v[i] / v[16 + i], 1), "count": int(v[16 + i])}
out["cycles_per_iteration"] = round(v[5] / max(v[7], 1), 1)
print(json.dumps(out))
