#!/usr/bin/env python
"""Generates tests/golden/*.npz with the cv2 wheel of this image (OpenCV 4.13.0) calling the SAME OpenCV routines
the reference calls (vins/estimator/feature_tracker.cpp:41-42,50-53,90,117-119,125-127,238,243), on the deterministic
synthetic input of vins_fast_b200/synth.py.  The fixtures pin the plain-C oracle (oracle/vf_oracle.c) on machines
without cv2 and on the GPU box.  Run from the repository root:   python tests/golden/make_golden.py
"""
import hashlib
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.camera import EUROC_CAM0, EUROC_CAM1, pinhole_for  # noqa: E402
from oracle.cv2_oracle import Cv2FeatureTracker, stdsort_perm  # noqa: E402
from vins_fast_b200.synth import StereoSequence  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def stages():
    """Stage-level known answers on a 160x200 crop (odd-sized variants included)."""
    seq = StereoSequence(0, 752, 480)
    _, l0, r0 = seq.frame(0)
    _, l1, _ = seq.frame(1)
    a, b, c = l0[40:200, 100:300].copy(), l1[40:200, 100:300].copy(), r0[40:200, 100:300].copy()
    out = {"crop_sha_a": sha(a), "crop_sha_b": sha(b), "crop_sha_c": sha(c)}
    for tag, img in (("even", a), ("odd", a[:159, :199].copy())):
        _, pyr = cv2.buildOpticalFlowPyramid(img, (21, 21), 3, withDerivatives=True)
        n = len(pyr) // 2
        out[f"pyr_{tag}_n"] = n
        for l in range(n):
            out[f"pyr_{tag}_{l}"] = np.ascontiguousarray(pyr[2 * l])        # the Python binding returns the unpadded ROI
            out[f"scharr_{tag}_{l}"] = np.ascontiguousarray(pyr[2 * l + 1])
    cv2.setUseOptimized(False)
    out["eig_noopt"] = cv2.cornerMinEigenVal(a, 3, ksize=3)
    cv2.setUseOptimized(True)
    out["eig_opt"] = cv2.cornerMinEigenVal(a, 3, ksize=3)
    out["gftt_60_10"] = cv2.goodFeaturesToTrack(a, 60, 0.01, 10).reshape(-1, 2)
    out["gftt_500_3"] = cv2.goodFeaturesToTrack(a, 500, 0.01, 3).reshape(-1, 2)
    mask = np.full(a.shape, 255, np.uint8)
    rng = np.random.default_rng(3)
    for _ in range(25):
        cv2.circle(mask, (int(rng.integers(0, 200)), int(rng.integers(0, 160))), int(rng.integers(0, 40)), 0, -1)
    out["mask"] = mask
    out["gftt_masked"] = cv2.goodFeaturesToTrack(a, 60, 0.01, 10, mask=mask).reshape(-1, 2)
    pts = cv2.goodFeaturesToTrack(a, 300, 0.01, 5).reshape(-1, 2)
    pts = (pts + rng.uniform(-0.5, 0.5, pts.shape)).astype(np.float32)
    pts = np.concatenate([pts, np.array([[0.3, 0.2], [199, 159], [2.5, 80], [197.7, 100.2], [-5, -5], [210, 30]], np.float32)])
    out["lk_pts"] = pts
    for tag, (x, y) in (("temporal", (a, b)), ("stereo", (a, c))):
        nxt, st, _ = cv2.calcOpticalFlowPyrLK(x, y, pts.reshape(-1, 1, 2), None, winSize=(21, 21), maxLevel=3)
        out[f"lk_{tag}_pts"], out[f"lk_{tag}_st"] = nxt.reshape(-1, 2), st.reshape(-1)
        rev, rst, _ = cv2.calcOpticalFlowPyrLK(y, x, nxt.copy(), pts.reshape(-1, 1, 2).copy(), winSize=(21, 21), maxLevel=1,
                                               criteria=(cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 30, 0.01),
                                               flags=cv2.OPTFLOW_USE_INITIAL_FLOW)
        out[f"lk_{tag}_rev_pts"], out[f"lk_{tag}_rev_st"] = rev.reshape(-1, 2), rst.reshape(-1)
    # std::sort permutations of libstdc++ for tracker-like track counts
    for n in (17, 150, 400):
        cnt = np.sort(rng.integers(1, 9, n)).astype(np.int32)[::-1].copy()
        cnt[rng.integers(0, n, n // 8 + 1)] += 2
        out[f"sort_cnt_{n}"], out[f"sort_perm_{n}"] = cnt, stdsort_perm(cnt)
    np.savez_compressed(os.path.join(OUT, "stages_cv2_4_13.npz"), **out)


def sequence(name, w, h, cam0, cam1, max_cnt, min_dist, flow_back, level, n_frames, seed):
    seq = StereoSequence(seed, w, h)
    tr = Cv2FeatureTracker(cam0, cam1, max_cnt, min_dist, flow_back, level)
    out = {"params": np.array([w, h, max_cnt, min_dist, flow_back, level, n_frames, seed], np.int64)}
    for k in range(n_frames):
        t, l, r = seq.frame(k)
        out[f"in_sha_{k}"] = np.frombuffer(bytes.fromhex(sha(l) + sha(r)), np.uint8)
        f = tr.track_image(t, l, r)
        for key, v in f.items():
            out[f"f{k}_{key}"] = v
    np.savez_compressed(os.path.join(OUT, f"sequence_{name}_cv2_4_13.npz"), **out)


if __name__ == "__main__":
    print("cv2", cv2.__version__)
    stages()
    sequence("euroc_mei", 752, 480, EUROC_CAM0, EUROC_CAM1, 150, 30, 1, 3, 8, 0)
    sequence("qvga_pinhole", 320, 240, pinhole_for(320, 240), pinhole_for(320, 240), 60, 20, 1, 3, 10, 5)
    sequence("qvga_noflowback", 320, 240, pinhole_for(320, 240), pinhole_for(320, 240), 40, 12, 0, 2, 6, 7)
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(OUT, f)))
