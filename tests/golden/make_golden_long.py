#!/usr/bin/env python
"""Long TrackImage sequences through the cv2 4.13.0 restatement of the reference (oracle/cv2_oracle.py: the same OpenCV
routines with the reference's arguments, vins/estimator/feature_tracker.cpp:27-217), kept as ONE SHA-256 PER FRAME over
the nine result arrays (tests/conftest.py::frame_digest) plus the feature counts -- long enough for track counts to
diversify and for features to churn (>= 100 frames).  Run from the repository root:
    python tests/golden/make_golden_long.py
"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import frame_digest  # noqa: E402
from oracle.camera import EUROC_CAM0, EUROC_CAM1, pinhole_for  # noqa: E402
from oracle.cv2_oracle import Cv2FeatureTracker  # noqa: E402
from vins_fast_b200.synth import StereoSequence  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def sequence(name, w, h, cam0, cam1, max_cnt, min_dist, flow_back, level, n_frames, seed):
    seq = StereoSequence(seed, w, h)
    tr = Cv2FeatureTracker(cam0, cam1, max_cnt, min_dist, flow_back, level)
    dig = np.zeros((n_frames, 32), np.uint8)
    counts = np.zeros((n_frames, 4), np.int32)   # n, n_right, max track count, landmark id after the frame
    for k in range(n_frames):
        t, l, r = seq.frame(k)
        f = tr.track_image(t, l, r)
        dig[k] = np.frombuffer(frame_digest(f), np.uint8)
        counts[k] = (len(f["ids"]), len(f["ids_right"]), int(f["track_cnt"].max(initial=0)), tr.landmark_id)
    np.savez_compressed(os.path.join(OUT, f"digest_{name}_cv2_4_13.npz"), digest=dig, counts=counts,
                        params=np.array([w, h, max_cnt, min_dist, flow_back, level, n_frames, seed], np.int64))
    print(name, "frames", n_frames, "features/frame", counts[:, 0].mean(), "stereo", counts[:, 1].mean(),
          "max track count", counts[:, 2].max(), "ids issued", counts[-1, 3])


if __name__ == "__main__":
    print("cv2", cv2.__version__)
    sequence("euroc_mei_120", 752, 480, EUROC_CAM0, EUROC_CAM1, 150, 30, 1, 3, 120, 0)
    sequence("720p_pinhole_100", 1280, 720, pinhole_for(1280, 720), pinhole_for(1280, 720), 400, 30, 1, 4, 100, 2)
