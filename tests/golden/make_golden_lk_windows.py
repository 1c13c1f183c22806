#!/usr/bin/env python
"""Generates tests/golden/lk_windows_cv2_4_13.npz: cv2.calcOpticalFlowPyrLK (OpenCV 4.13.0 wheel of this image) with window
sizes OTHER than the reference's literal cv::Size(21, 21) (vins/estimator/feature_tracker.cpp:42,51,119,127) -- the north
star keeps the window size as a parameter -- on the crops of make_golden.py, plus two short whole-TrackImage sequences of the
cv2 restatement with lk_win 15 and 31.  OpenCV's float summation order depends on the width (8-wide SIMD groups + scalar
tail), so every width pins its own chain layout of oracle/vf_oracle.c.  Run from the repository root:
    python tests/golden/make_golden_lk_windows.py
"""
import hashlib
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.camera import pinhole_for  # noqa: E402
from oracle.cv2_oracle import Cv2FeatureTracker  # noqa: E402
from vins_fast_b200.synth import StereoSequence  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
WINDOWS = (5, 7, 9, 11, 13, 15, 17, 19, 23, 25, 27, 29, 31)
SEQUENCES = (("win15", 15, 3, 6, 11), ("win31", 31, 2, 6, 12))     # name, lk_win, maxLevel, frames, seed


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    print("cv2", cv2.__version__)
    seq = StereoSequence(0, 752, 480)
    _, l0, r0 = seq.frame(0)
    _, l1, _ = seq.frame(1)
    a, b = l0[40:200, 100:300].copy(), l1[40:200, 100:300].copy()
    rng = np.random.default_rng(9)
    pts = cv2.goodFeaturesToTrack(a, 120, 0.01, 7).reshape(-1, 2)
    pts = (pts + rng.uniform(-0.5, 0.5, pts.shape)).astype(np.float32)
    pts = np.concatenate([pts, np.array([[0.3, 0.2], [199, 159], [2.5, 80], [197.7, 100.2], [-5, -5], [210, 30], [-40, 50]], np.float32)])
    out = {"crop_sha_a": sha(a), "crop_sha_b": sha(b), "pts": pts, "windows": np.array(WINDOWS, np.int32)}
    for win in WINDOWS:
        for level in (0, 3):
            nxt, st, _ = cv2.calcOpticalFlowPyrLK(a, b, pts.reshape(-1, 1, 2), None, winSize=(win, win), maxLevel=level)
            out[f"w{win}_l{level}_pts"], out[f"w{win}_l{level}_st"] = nxt.reshape(-1, 2), st.reshape(-1)
        rev, rst, _ = cv2.calcOpticalFlowPyrLK(b, a, nxt.copy(), pts.reshape(-1, 1, 2).copy(), winSize=(win, win), maxLevel=1,
                                               criteria=(cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 30, 0.01),
                                               flags=cv2.OPTFLOW_USE_INITIAL_FLOW)
        out[f"w{win}_rev_pts"], out[f"w{win}_rev_st"] = rev.reshape(-1, 2), rst.reshape(-1)
    for name, win, level, n_frames, seed in SEQUENCES:
        w, h, max_cnt, min_dist = 320, 240, 60, 20
        cam = pinhole_for(w, h)
        s = StereoSequence(seed, w, h)
        tr = Cv2FeatureTracker(cam, cam, max_cnt, min_dist, 1, level, lk_win=win)
        out[f"{name}_params"] = np.array([w, h, max_cnt, min_dist, 1, level, n_frames, seed, win], np.int64)
        for k in range(n_frames):
            t, l, r = s.frame(k)
            f = tr.track_image(t, l, r)
            for key, v in f.items():
                out[f"{name}_f{k}_{key}"] = v
    path = os.path.join(OUT, "lk_windows_cv2_4_13.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path))


if __name__ == "__main__":
    main()
