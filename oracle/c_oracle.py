"""TEST INFRASTRUCTURE (oracle) -- ctypes binding of oracle/vf_oracle.c (plain-C CPU restatement).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from .camera import Camera

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libvf_oracle.so")
_lib = None

VFO_MAX_LEVELS = 8


class _Pyr(C.Structure):
    _fields_ = [("n_levels", C.c_int), ("w", C.c_int * VFO_MAX_LEVELS), ("h", C.c_int * VFO_MAX_LEVELS),
                ("img", C.c_void_p * VFO_MAX_LEVELS)]


class _Cam(C.Structure):
    _fields_ = [("model", C.c_int)] + [(n, C.c_double) for n in ("xi", "k1", "k2", "p1", "p2", "fx", "fy", "cx", "cy")]


class Feature(C.Structure):
    _fields_ = [("id", C.c_int32), ("track_cnt", C.c_int32), ("has_right", C.c_int32)] + \
               [(n, C.c_float) for n in ("x", "y", "u", "v", "vx", "vy", "xr", "yr", "ur", "vr", "vxr", "vyr")]


FEATURE_DTYPE = np.dtype([("id", "<i4"), ("track_cnt", "<i4"), ("has_right", "<i4")] +
                         [(n, "<f4") for n in ("x", "y", "u", "v", "vx", "vy", "xr", "yr", "ur", "vr", "vxr", "vyr")])


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("vf_oracle.c", "vf_oracle.h", "stdsort_helper.cpp", "Makefile")]
    stale = (not os.path.exists(_SO)) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs)
    if force or stale:
        subprocess.check_call(["make", "-s", "-C", _HERE], stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.vfo_tracker_create.restype = C.c_void_p
        _lib.vfo_tracker_create.argtypes = [C.c_int] * 7 + [C.c_double, C.c_void_p, C.c_void_p, C.c_int]
        _lib.vfo_tracker_destroy.argtypes = [C.c_void_p]
        _lib.vfo_tracker_track.restype = C.c_int
        _lib.vfo_tracker_track.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.vfo_good_features.restype = C.c_int
        _lib.vfo_good_features.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int,
                                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.vfo_lk.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                C.c_int, C.c_int, C.c_int, C.c_double, C.c_float, C.c_int, C.c_void_p]
        _lib.vfo_pyramid_build.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
        _lib.vfo_pyramid_free.argtypes = [C.c_void_p]
        _lib.vfo_pyrdown_u8.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        _lib.vfo_scharr_s16.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        _lib.vfo_min_eigen_val.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        _lib.vfo_mask_disc.argtypes = [C.c_void_p] + [C.c_int] * 5
        _lib.vfo_undistort.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.c_void_p]
    return _lib


def _cam(c: Camera) -> _Cam:
    return _Cam(c.model, c.xi, c.k1, c.k2, c.p1, c.p2, c.fx, c.fy, c.cx, c.cy)


def _u8(img) -> np.ndarray:
    a = np.ascontiguousarray(img, dtype=np.uint8)
    assert a.ndim == 2
    return a


def pyrdown(img) -> np.ndarray:
    a = _u8(img)
    h, w = a.shape
    out = np.empty(((h + 1) // 2, (w + 1) // 2), np.uint8)
    lib().vfo_pyrdown_u8(a.ctypes.data, w, h, out.ctypes.data)
    return out


class Pyramid:
    def __init__(self, img, max_level: int = 3, win: int = 21):
        a = _u8(img)
        self._p = _Pyr()
        lib().vfo_pyramid_build(C.byref(self._p), a.ctypes.data, a.shape[1], a.shape[0], max_level, win)

    @property
    def n_levels(self):
        return self._p.n_levels

    def level(self, l: int) -> np.ndarray:
        w, h = self._p.w[l], self._p.h[l]
        buf = (C.c_uint8 * (w * h)).from_address(self._p.img[l])
        return np.frombuffer(buf, np.uint8).reshape(h, w).copy()

    def __del__(self):
        try:
            lib().vfo_pyramid_free(C.byref(self._p))
        except Exception:
            pass


def scharr(img) -> np.ndarray:
    a = _u8(img)
    h, w = a.shape
    out = np.empty((h, w, 2), np.int16)
    lib().vfo_scharr_s16(a.ctypes.data, w, h, out.ctypes.data)
    return out


def lk(prev: Pyramid, nxt: Pyramid, prev_pts, max_level=3, win=21, init=None, max_count=30, eps=0.01,
       min_eig=1e-4, accum=0, return_iters=False):
    p = np.ascontiguousarray(prev_pts, np.float32).reshape(-1, 2)
    n = len(p)
    q = np.zeros((n, 2), np.float32) if init is None else np.ascontiguousarray(init, np.float32).reshape(-1, 2).copy()
    st = np.zeros(n, np.uint8)
    it = np.zeros(n, np.int32)
    lib().vfo_lk(C.byref(prev._p), C.byref(nxt._p), p.ctypes.data, q.ctypes.data, st.ctypes.data, n, max_level, win,
                 0 if init is None else 1, max_count, eps, min_eig, accum, it.ctypes.data)
    return (q, st, it) if return_iters else (q, st)


def min_eigen_val(img) -> np.ndarray:
    a = _u8(img)
    h, w = a.shape
    out = np.empty((h, w), np.float32)
    lib().vfo_min_eigen_val(a.ctypes.data, w, h, out.ctypes.data)
    return out


def good_features(img, max_corners, quality, min_dist, mask=None, eig=None):
    a = _u8(img)
    h, w = a.shape
    cap = max_corners if max_corners > 0 else w * h
    out = np.zeros((cap, 2), np.float32)
    m = None if mask is None else _u8(mask)
    e = None if eig is None else np.ascontiguousarray(eig, np.float32)
    nc = C.c_int(0)
    n = lib().vfo_good_features(a.ctypes.data, w, h, max_corners, quality, min_dist,
                                None if m is None else m.ctypes.data, None if e is None else e.ctypes.data,
                                out.ctypes.data, C.byref(nc))
    return out[:n].copy(), nc.value


def mask_disc(mask: np.ndarray, cx: int, cy: int, r: int):
    assert mask.dtype == np.uint8 and mask.flags.c_contiguous
    lib().vfo_mask_disc(mask.ctypes.data, mask.shape[1], mask.shape[0], cx, cy, r)


def undistort(cam: Camera, u: float, v: float):
    c = _cam(cam)
    x, y = C.c_double(), C.c_double()
    lib().vfo_undistort(C.byref(c), float(u), float(v), C.byref(x), C.byref(y))
    return x.value, y.value


class CFeatureTracker:
    """Whole TrackImage in plain C (same interface as oracle.cv2_oracle.Cv2FeatureTracker.track_image)."""

    def __init__(self, cam0: Camera, cam1: Camera, width: int, height: int, max_cnt=150, min_dist=30, flow_back=1,
                 lk_max_level=3, lk_win=21, quality=0.01, lk_accum=0):
        self.w, self.h, self.max_cnt = width, height, max_cnt
        c0, c1 = _cam(cam0), _cam(cam1)
        self._t = lib().vfo_tracker_create(width, height, max_cnt, min_dist, flow_back, lk_max_level, lk_win,
                                           quality, C.byref(c0), C.byref(c1), lk_accum)
        self._out = np.zeros(max(max_cnt, 1), FEATURE_DTYPE)

    def track_image(self, t: float, img0, img1) -> dict:
        a, b = _u8(img0), _u8(img1)
        assert a.shape == (self.h, self.w) and b.shape == (self.h, self.w)
        n = lib().vfo_tracker_track(self._t, float(t), a.ctypes.data, b.ctypes.data, self._out.ctypes.data)
        return records_to_frame(self._out[:n].copy())

    def __del__(self):
        try:
            lib().vfo_tracker_destroy(self._t)
        except Exception:
            pass


def records_to_frame(rec: np.ndarray) -> dict:
    """vf_feature[] -> the dict layout of Cv2FeatureTracker.track_image."""
    r = rec[rec["has_right"] != 0]
    return {
        "ids": rec["id"].astype(np.int32), "track_cnt": rec["track_cnt"].astype(np.int32),
        "uv": np.stack([rec["u"], rec["v"]], 1).astype(np.float32), "un": np.stack([rec["x"], rec["y"]], 1).astype(np.float32),
        "vel": np.stack([rec["vx"], rec["vy"]], 1).astype(np.float32),
        "ids_right": r["id"].astype(np.int32), "uv_right": np.stack([r["ur"], r["vr"]], 1).astype(np.float32),
        "un_right": np.stack([r["xr"], r["yr"]], 1).astype(np.float32),
        "vel_right": np.stack([r["vxr"], r["vyr"]], 1).astype(np.float32),
    }
