#!/usr/bin/env python
"""Per-CTA duration of the 4-warp LK kernels against the per-feature iteration counters: base + us per iteration + us per
iteration that ran the ordered float chains (least-squares fit), and the five longest CTAs of a few frames.
usage (on a B200): python -m vins_fast_b200.build --variant lktiming -DVF_LK_TIMING
                   VF_LIB=vins_fast_b200/libvinsfast_b200_lktiming.so VF_LK_KERNEL=1 python profiles/lk_cta_cold.py"""
import ctypes as C, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vins_fast_b200 as vf
from vins_fast_b200 import _lib
from vins_fast_b200.calib import EUROC_CAM0, EUROC_CAM1
from vins_fast_b200.synth import StereoSequence
seq = StereoSequence(0, 752, 480)
tr = vf.FeatureTracker(vf.make_config(752, 480, 150, 30, 1, EUROC_CAM0, EUROC_CAM1))
lib = vf.load()
buf = (C.c_ulonglong * (2 * 256 * 4))()
rows = []
for k in range(40):
    t, l, r = seq.frame(k)
    tr.track(t, l, r)
    if k < 10:
        continue
    lib.vf_debug_lk_cta_times(buf)
    a = np.array(list(buf), np.int64).reshape(2, 256, 4)
    for kern in (0, 1):
        it = tr.debug(_lib.DBG_LK_ITERS, kern, dtype=np.int32)
        n = len(it)
        v = a[kern, :n]
        dur = (v[:, 1] - v[:, 0]) / 1e3
        iters, fast = it & 0xffff, it >> 16
        cold = iters - fast
        ok = dur > 0
        if ok.sum() < 10:
            continue
        A = np.stack([np.ones(ok.sum()), iters[ok], cold[ok]], 1)
        coef = np.linalg.lstsq(A, dur[ok], rcond=None)[0]
        top = np.argsort(-dur)[:5]
        rows.append({"kern": kern, "frame": k, "n": int(n), "span": float(dur.max()), "mean": float(dur[ok].mean()),
                     "fit_base_us": float(coef[0]), "fit_us_per_iter": float(coef[1]), "fit_us_per_cold": float(coef[2]),
                     "iters_mean": float(iters.mean()), "cold_mean": float(cold.mean()),
                     "top5": [(round(float(dur[i]), 1), int(iters[i]), int(cold[i])) for i in top]})
for kern in (0, 1):
    rr = [x for x in rows if x["kern"] == kern]
    print(json.dumps({"kern": kern, **{k: round(float(np.mean([x[k] for x in rr])), 2) for k in ("span", "mean", "fit_base_us", "fit_us_per_iter", "fit_us_per_cold", "iters_mean", "cold_mean")}}))
    for x in rr[:6]:
        print("   ", x["frame"], x["top5"])
