"""TEST INFRASTRUCTURE (oracle) -- the reference's TrackImage restated over the real OpenCV routines.

`FeatureTracker::TrackImage` (vins/estimator/feature_tracker.cpp:27-217) is glue around OpenCV calls;
OpenCV itself is a third-party dependency that is NOT vendored in the reference
(vins/cmake/packages.cmake:11 -> find_package(OpenCV 3 REQUIRED); no exact version pinned).  This file
restates the glue line by line and calls the *same* OpenCV routines (through the `cv2` 4.13.0 wheel of
this image, the closest obtainable stand-in for the author's OpenCV 3.x C++ build) with the reference's
exact arguments:

  feature_tracker.cpp:20-25    InBorder                   -> _in_border (cvRound = round-half-even = np.rint)
  feature_tracker.cpp:41-42    calcOpticalFlowPyrLK fwd   -> winSize (21,21), maxLevel 3, default criteria 30/0.01
  feature_tracker.cpp:50-53    calcOpticalFlowPyrLK rev   -> maxLevel 1, criteria (COUNT+EPS,30,0.01), USE_INITIAL_FLOW
  feature_tracker.cpp:55-59    status &= fwd & rev & |prev - reverse| <= 0.5
  feature_tracker.cpp:63-71    InBorder + ReduceVector
  feature_tracker.cpp:76-79    ++track_count
  feature_tracker.cpp:219-246  SetFeatureExtractorMask (std::sort through oracle/stdsort_helper.cpp + cv2.circle)
  feature_tracker.cpp:85-101   goodFeaturesToTrack(cur, n, 0.01, MIN_DIST, mask), new ids from landmark_id_++
  feature_tracker.cpp:103-104  RemoveDistortion + ComputeKpsVelocity (oracle/camera.py, f32 casts as in C++)
  feature_tracker.cpp:117-141  stereo L->R, R->L, status, ids_right_, right undistort + velocity
  feature_tracker.cpp:158-168  roll prev <- cur
  feature_tracker.cpp:173-213  featureFrame: {id: [(0,[x,y,1,u,v,vx,vy]), (1,[...])]}

PARITY UNPINNED: the reference holds no golden vectors, fixtures or assertions for this
path, so the oracle is anchored on OpenCV itself.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference leg may import this module; the product path never does.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import time
from typing import Optional

import numpy as np

try:  # cv2 is present in this image (4.13.0); anything without it falls back to the C restatement
    import cv2
except Exception:  # pragma: no cover
    cv2 = None

from .camera import Camera

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, "_build")
_stdsort_lib = None


def build_stdsort_helper(force: bool = False) -> str:
    os.makedirs(_BUILD, exist_ok=True)
    so = os.path.join(_BUILD, "libvf_stdsort.so")
    src = os.path.join(_HERE, "stdsort_helper.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["g++", "-O2", "-std=c++14", "-shared", "-fPIC", "-o", so, src])
    return so


def stdsort_perm(track_count) -> np.ndarray:
    """Permutation of libstdc++ std::sort with the reference's comparator (feature_tracker.cpp:228-231)."""
    global _stdsort_lib
    if _stdsort_lib is None:
        _stdsort_lib = ctypes.CDLL(build_stdsort_helper())
        _stdsort_lib.vf_oracle_stdsort_perm.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
    tc = np.ascontiguousarray(track_count, dtype=np.int32)
    perm = np.empty(len(tc), dtype=np.int32)
    if len(tc):
        _stdsort_lib.vf_oracle_stdsort_perm(tc.ctypes.data, len(tc), perm.ctypes.data)
    return perm


def _in_border(pts, rows: int, cols: int) -> np.ndarray:
    """InBorder (feature_tracker.cpp:20-25) for an [N,2] float32 array: cvRound = round-half-even = np.rint."""
    p = np.asarray(pts, np.float32).reshape(-1, 2)
    x = np.rint(p[:, 0]).astype(np.int64)
    y = np.rint(p[:, 1]).astype(np.int64)
    return (1 <= x) & (x < cols - 1) & (1 <= y) & (y < rows - 1)


def _pt_distance(a, b) -> np.ndarray:
    """PtDistance (feature_tracker.cpp:457-462) row-wise: float subtraction, then widened to double."""
    a = np.asarray(a, np.float32).reshape(-1, 2)
    b = np.asarray(b, np.float32).reshape(-1, 2)
    dx = (a[:, 0] - b[:, 0]).astype(np.float64)
    dy = (a[:, 1] - b[:, 1]).astype(np.float64)
    return np.sqrt(dx * dx + dy * dy)


class Cv2FeatureTracker:
    """Line-by-line restatement of estimator::FeatureTracker over cv2.

    The glue between the OpenCV calls is vectorised with numpy (same float32 / float64 expressions, element-wise, so the
    results are bit-identical to a per-point loop): the C++ reference has no interpreter between its OpenCV calls, so a
    CPU baseline timed through this class must not carry one either.  `cv2_seconds` accumulates the time spent INSIDE the
    OpenCV routines (4x calcOpticalFlowPyrLK, goodFeaturesToTrack, the circles of SetFeatureExtractorMask);
    `calls` counts the frames, so bench.py can report "cv2 calls only" beside the inclusive figure.
    """

    LK_WIN = (21, 21)            # the reference's literal cv::Size(21, 21) (feature_tracker.cpp:42,51,119,127)

    def __init__(self, cam0: Camera, cam1: Camera, max_cnt: int = 150, min_dist: int = 30,
                 flow_back: int = 1, lk_max_level: int = 3, quality: float = 0.01, keep_debug: bool = True, lk_win: int = 21):
        if cv2 is None:
            raise RuntimeError("cv2 is not importable; use oracle.c_oracle instead")
        self.LK_WIN = (int(lk_win), int(lk_win))     # the north star keeps the window size as a parameter
        self.cams = [cam0, cam1]
        self.MAX_CNT, self.MIN_DIST, self.FLOW_BACK = int(max_cnt), int(min_dist), int(flow_back)
        self.lk_max_level, self.quality = int(lk_max_level), float(quality)
        self.keep_debug = bool(keep_debug)
        self.landmark_id = 0
        self.prev_time = 0.0
        self.cur_time = 0.0
        self.prev_img: Optional[np.ndarray] = None
        self.prev_kps = np.zeros((0, 2), np.float32)
        self.ids = np.zeros(0, np.int64)
        self.track_count = np.zeros(0, np.int32)
        self.prev_left_un: dict = {}      # id -> row of prev_left_un_pts (ComputeKpsVelocity's previous map)
        self.prev_right_un: dict = {}
        self.prev_left_un_pts = np.zeros((0, 2), np.float32)
        self.prev_right_un_pts = np.zeros((0, 2), np.float32)
        self.debug: dict = {}
        self.cv2_seconds = 0.0
        self.calls = 0

    # -- helpers ---------------------------------------------------------------------------------------
    def _lk(self, a, b, pts, max_level, init=None):
        p = np.ascontiguousarray(pts, np.float32).reshape(-1, 1, 2)
        if len(p) == 0:
            return np.zeros((0, 2), np.float32), np.zeros((0,), np.uint8)
        t0 = time.perf_counter()
        if init is None:
            nxt, st, _ = cv2.calcOpticalFlowPyrLK(a, b, p, None, winSize=self.LK_WIN, maxLevel=max_level)
        else:
            q = np.ascontiguousarray(init, np.float32).reshape(-1, 1, 2).copy()
            nxt, st, _ = cv2.calcOpticalFlowPyrLK(
                a, b, p, q, winSize=self.LK_WIN, maxLevel=max_level,
                criteria=(cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 30, 0.01),
                flags=cv2.OPTFLOW_USE_INITIAL_FLOW)
        self.cv2_seconds += time.perf_counter() - t0
        return nxt.reshape(-1, 2), st.reshape(-1)

    def _undistort(self, pts, cam: Camera) -> np.ndarray:
        return cam.undistort_many(np.asarray(pts, np.float32).reshape(-1, 2))

    def _velocity(self, ids, un, prev_map, prev_pts):
        """ComputeKpsVelocity (feature_tracker.cpp:323-363): v = (f32 cur - f32 prev) / (double) dt for ids found in the
        previous map, else 0; the current map keeps the FIRST point seen for an id (std::map::insert)."""
        ids_l = [int(v) for v in ids]
        cur_map = {}
        for i, fid in enumerate(ids_l):
            cur_map.setdefault(fid, i)
        vel = np.zeros((len(ids_l), 2), np.float32)
        if prev_map and len(ids_l):
            dt = self.cur_time - self.prev_time
            hit = np.fromiter((prev_map.get(f, -1) for f in ids_l), np.int64, len(ids_l))
            sel = hit >= 0
            d = (un[sel] - prev_pts[hit[sel]]).astype(np.float64)      # float32 subtraction, then widened
            with np.errstate(divide="ignore", invalid="ignore"):
                vel[sel] = (d / dt).astype(np.float32)
        return vel, cur_map

    def _set_mask(self, img_shape, cur_kps):
        rows, cols = img_shape
        mask = np.full((rows, cols), 255, np.uint8)
        perm = stdsort_perm(self.track_count)
        centres = np.rint(np.asarray(cur_kps, np.float32).reshape(-1, 2)).astype(np.int64)
        kept = []
        r = self.MIN_DIST
        t0 = time.perf_counter()
        for j in perm.tolist():
            cx, cy = int(centres[j, 0]), int(centres[j, 1])
            if mask[cy, cx] == 255:
                kept.append(j)
                cv2.circle(mask, (cx, cy), r, 0, -1)
        self.cv2_seconds += time.perf_counter() - t0                       # the loop is the reference's own (its C++ twin is free)
        kept = np.asarray(kept, np.int64)
        self.ids, self.track_count = self.ids[kept], self.track_count[kept]
        return np.asarray(cur_kps, np.float32).reshape(-1, 2)[kept], mask, perm

    # -- TrackImage ------------------------------------------------------------------------------------
    def track_image(self, t: float, img0: np.ndarray, img1: np.ndarray) -> dict:
        assert img0 is not None and img1 is not None and img0.size and img1.size
        dbg = self.debug = {}
        keep_dbg = self.keep_debug
        self.calls += 1
        self.cur_time = float(t)
        rows, cols = img0.shape
        cur_kps = np.zeros((0, 2), np.float32)

        if len(self.prev_kps):
            cur_kps, status = self._lk(self.prev_img, img0, self.prev_kps, self.lk_max_level)
            status = status != 0
            if keep_dbg:
                dbg["fwd_pts"], dbg["fwd_status"] = cur_kps.copy(), status.astype(np.uint8)
            if self.FLOW_BACK:
                rev, rstatus = self._lk(img0, self.prev_img, cur_kps, 1, init=self.prev_kps)
                if keep_dbg:
                    dbg["rev_pts"], dbg["rev_status"] = rev.copy(), rstatus.copy()
                status = status & (rstatus != 0) & (_pt_distance(self.prev_kps, rev) <= 0.5)
            status = status & _in_border(cur_kps, rows, cols)
            if keep_dbg:
                dbg["temporal_status"] = status.astype(np.uint8)
            cur_kps = cur_kps[status]
            self.ids = self.ids[status]
            self.track_count = self.track_count[status]

        self.track_count = self.track_count + 1

        cur_kps, mask, perm = self._set_mask((rows, cols), cur_kps)
        if keep_dbg:
            dbg["sort_perm"], dbg["mask"] = perm, mask
            dbg["n_kept"] = len(cur_kps)
        n_new = self.MAX_CNT - len(cur_kps)
        new_pts = np.zeros((0, 2), np.float32)
        if n_new > 0:
            t0 = time.perf_counter()
            r = cv2.goodFeaturesToTrack(img0, n_new, self.quality, self.MIN_DIST, mask=mask)
            self.cv2_seconds += time.perf_counter() - t0
            if r is not None:
                new_pts = r.reshape(-1, 2).astype(np.float32)
        if keep_dbg:
            dbg["new_pts"] = new_pts.copy()
        k = len(new_pts)
        if k:
            self.track_count = np.concatenate([self.track_count, np.ones(k, np.int32)])
            self.ids = np.concatenate([self.ids, (self.landmark_id + np.arange(k, dtype=np.int64)) & 0xFFFFFFFF])
            self.landmark_id = (self.landmark_id + k) & 0xFFFFFFFF
        cur_kps = np.concatenate([cur_kps, new_pts], axis=0).astype(np.float32)

        cur_un = self._undistort(cur_kps, self.cams[0])
        vel, cur_left_map = self._velocity(self.ids, cur_un, self.prev_left_un, self.prev_left_un_pts)

        # stereo
        right_kps, status = self._lk(img0, img1, cur_kps, self.lk_max_level)
        status = status != 0
        if keep_dbg:
            dbg["lr_pts"], dbg["lr_status"] = right_kps.copy(), status.astype(np.uint8)
        if self.FLOW_BACK and len(cur_kps):
            rev_left, rstatus = self._lk(img1, img0, right_kps, self.lk_max_level)
            if keep_dbg:
                dbg["rl_pts"], dbg["rl_status"] = rev_left.copy(), rstatus.copy()
            status = status & (rstatus != 0) & _in_border(right_kps, rows, cols) & (_pt_distance(cur_kps, rev_left) <= 0.5)
        if keep_dbg:
            dbg["stereo_status"] = status.astype(np.uint8)
        ids_right = self.ids[status]
        right_kps = right_kps[status]
        right_un = self._undistort(right_kps, self.cams[1])
        rvel, cur_right_map = self._velocity(ids_right, right_un, self.prev_right_un, self.prev_right_un_pts)

        # roll state
        self.prev_img = img0
        self.prev_kps = cur_kps
        self.prev_left_un, self.prev_left_un_pts = cur_left_map, cur_un
        self.prev_right_un, self.prev_right_un_pts = cur_right_map, right_un
        self.prev_time = self.cur_time

        frame = {
            "ids": self.ids.astype(np.int32), "track_cnt": self.track_count.astype(np.int32),
            "uv": cur_kps.copy(), "un": cur_un, "vel": vel,
            "ids_right": ids_right.astype(np.int32), "uv_right": right_kps.copy(),
            "un_right": right_un, "vel_right": rvel,
        }
        return frame

    @staticmethod
    def feature_frame(frame: dict) -> dict:
        """featureFrame map (feature_tracker.cpp:173-213): id -> [(cam, [x,y,z,u,v,vx,vy]) ...]."""
        ff: dict = {}
        for i, fid in enumerate(frame["ids"]):
            v = [frame["un"][i, 0], frame["un"][i, 1], 1.0, frame["uv"][i, 0], frame["uv"][i, 1],
                 frame["vel"][i, 0], frame["vel"][i, 1]]
            ff.setdefault(int(fid), []).append((0, np.array(v, np.float64)))
        for i, fid in enumerate(frame["ids_right"]):
            v = [frame["un_right"][i, 0], frame["un_right"][i, 1], 1.0, frame["uv_right"][i, 0],
                 frame["uv_right"][i, 1], frame["vel_right"][i, 0], frame["vel_right"][i, 1]]
            ff.setdefault(int(fid), []).append((1, np.array(v, np.float64)))
        return ff
