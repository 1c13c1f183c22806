"""Stand-alone stage operators (host numpy in, host numpy out), each one mirroring the OpenCV / camodocal
routine the reference reaches from FeatureTracker::TrackImage.  They run the same sm_100a kernels as the
tracker and exist for the parity tests; there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import check, load
from .tracker import _img, camera_struct


def build_pyramid(img, max_level: int = 3, device: int = -1):
    """cv::buildOpticalFlowPyramid(img, Size(21,21), max_level) -> list of unpadded uint8 levels."""
    a = np.ascontiguousarray(_img(img))
    h, w = a.shape
    buf = np.zeros(2 * w * h, np.uint8)
    ws = (C.c_int32 * 8)()
    hs = (C.c_int32 * 8)()
    n = C.c_int32(0)
    check(load().vf_op_build_pyramid(a.ctypes.data, w, h, max_level, device, buf.ctypes.data, ws, hs, C.byref(n)))
    out, off = [], 0
    for l in range(n.value):
        out.append(buf[off: off + ws[l] * hs[l]].reshape(hs[l], ws[l]).copy())
        off += ws[l] * hs[l]
    return out


def scharr(img, device: int = -1) -> np.ndarray:
    """calcScharrDeriv -> int16 (h, w, 2) = (Ix, Iy)."""
    a = np.ascontiguousarray(_img(img))
    h, w = a.shape
    out = np.zeros((h, w, 2), np.int16)
    check(load().vf_op_scharr(a.ctypes.data, w, h, device, out.ctypes.data))
    return out


def lk(prev, nxt, prev_pts, max_level: int = 3, init=None, device: int = -1, return_iters: bool = False, win: int = 21):
    """cv::calcOpticalFlowPyrLK(prev, next, prev_pts, init, winSize=(win,win), maxLevel, criteria 30/0.01,
    flags = USE_INITIAL_FLOW if init is given) -> (next_pts, status).  win: odd, 5..31 (21 = the reference's)."""
    a, b = np.ascontiguousarray(_img(prev)), np.ascontiguousarray(_img(nxt))
    assert a.shape == b.shape
    h, w = a.shape
    p = np.ascontiguousarray(prev_pts, np.float32).reshape(-1, 2)
    n = len(p)
    q = np.zeros((n, 2), np.float32) if init is None else np.ascontiguousarray(init, np.float32).reshape(-1, 2).copy()
    st = np.zeros(n, np.uint8)
    it = np.zeros(n, np.int32)
    check(load().vf_op_lk_win(a.ctypes.data, b.ctypes.data, w, h, p.ctypes.data, q.ctypes.data, st.ctypes.data, n, max_level,
                              int(win), 0 if init is None else 1, device, it.ctypes.data))
    return (q, st, it) if return_iters else (q, st)


def set_response_kernel(mode: int) -> None:
    """Parity tests: 0 = pick the corner-response kernel by frame size, 1 = tiles, 2 = strips (identical results)."""
    check(load().vf_debug_set_response_kernel(int(mode)))


def set_lk_kernel(mode: int) -> None:
    """A/B and parity tests: 0 = warp-level passes scheduled by launch size (default), 1 = the block-barrier kernels of vf_lk.cu,
    2 = one warp per feature, 3 = one 4-warp CTA per feature with warp-level passes (identical results)."""
    check(load().vf_debug_set_lk_kernel(int(mode)))


def set_pyramid_kernel(mode: int) -> None:
    """Parity tests: 0 = pick the pyramid kernel by size, 1 = fused multi-level kernel, 2 = streaming TMA kernel (identical results)."""
    check(load().vf_debug_set_pyramid_kernel(int(mode)))


def min_eigen(img, device: int = -1) -> np.ndarray:
    """cv::cornerMinEigenVal(img, 3, ksize=3) -> float32 (h, w)."""
    a = np.ascontiguousarray(_img(img))
    h, w = a.shape
    out = np.zeros((h, w), np.float32)
    check(load().vf_op_min_eigen(a.ctypes.data, w, h, device, out.ctypes.data))
    return out


def good_features(img, max_corners: int, quality: float, min_dist: int, mask=None, device: int = -1) -> np.ndarray:
    """cv::goodFeaturesToTrack(img, max_corners, quality, min_dist, mask) -> float32 (n, 2)."""
    a = np.ascontiguousarray(_img(img))
    h, w = a.shape
    m = None if mask is None else np.ascontiguousarray(_img(mask))
    out = np.zeros((max_corners, 2), np.float32)
    n = C.c_int32(0)
    check(load().vf_op_good_features(a.ctypes.data, w, h, max_corners, quality, min_dist,
                                     None if m is None else m.ctypes.data, device, out.ctypes.data, C.byref(n)))
    return out[: n.value].copy()


def set_mask(pts, track_cnt, width: int, height: int, min_dist: int, device: int = -1):
    """SetFeatureExtractorMask (feature_tracker.cpp:219-246) -> (std::sort permutation, kept indices in order, mask)."""
    p = np.ascontiguousarray(pts, np.float32).reshape(-1, 2)
    c = np.ascontiguousarray(track_cnt, np.int32)
    n = len(p)
    perm = np.zeros(max(n, 1), np.int32)
    keep = np.zeros(max(n, 1), np.int32)
    nk = C.c_int32(0)
    mask = np.zeros((height, width), np.uint8)
    check(load().vf_op_set_mask(p.ctypes.data, c.ctypes.data, n, width, height, min_dist, device, perm.ctypes.data,
                                keep.ctypes.data, C.byref(nk), mask.ctypes.data))
    return perm[:n].copy(), keep[: nk.value].copy(), mask


def undistort(cam, uv, device: int = -1) -> np.ndarray:
    """RemoveDistortion (feature_tracker.cpp:310-321): liftProjective + /z -> float32 (n, 2)."""
    c = camera_struct(cam)
    p = np.ascontiguousarray(uv, np.float32).reshape(-1, 2)
    out = np.zeros_like(p)
    check(load().vf_op_undistort(C.byref(c), p.ctypes.data, len(p), device, out.ctypes.data))
    return out


# ---- landmark geometry -----------------------------------------------------------------------------
def _window(Ps, Rs, tic, ric):
    from ._lib import VfWindow
    Ps = np.ascontiguousarray(Ps, np.float64).reshape(-1, 3)
    Rs = np.ascontiguousarray(Rs, np.float64).reshape(-1, 9)
    assert len(Ps) == len(Rs)
    w = VfWindow()
    w.window, w.Ps, w.Rs = len(Ps), Ps.ctypes.data, Rs.ctypes.data
    t = np.ascontiguousarray(tic, np.float64).reshape(2, 3)
    r = np.ascontiguousarray(ric, np.float64).reshape(2, 9)
    for c in range(2):
        for k in range(3):
            w.tic[c][k] = t[c, k]
        for k in range(9):
            w.ric[c][k] = r[c, k]
    return w, (Ps, Rs)          # keep the arrays alive while the struct points at them


def _table(start_frame, n_obs, obs_offset, obs):
    sf = np.ascontiguousarray(start_frame, np.int32)
    no = np.ascontiguousarray(n_obs, np.int32)
    oo = np.ascontiguousarray(obs_offset, np.int32)
    ob = np.ascontiguousarray(obs, np.float64).reshape(-1, 7)
    return sf, no, oo, ob


def triangulate_one_point(pose0, pose1, point0, point1, device: int = -1) -> np.ndarray:
    """FeatureManager::TriangulateOnePoint (feature_manager.cpp:277-293) for n landmarks -> float64 (n, 3)."""
    a = np.ascontiguousarray(pose0, np.float64).reshape(-1, 12)
    b = np.ascontiguousarray(pose1, np.float64).reshape(-1, 12)
    p0 = np.ascontiguousarray(point0, np.float64).reshape(-1, 2)
    p1 = np.ascontiguousarray(point1, np.float64).reshape(-1, 2)
    out = np.zeros((len(a), 3), np.float64)
    check(load().vf_op_triangulate_one_point(a.ctypes.data, b.ctypes.data, p0.ctypes.data, p1.ctypes.data, len(a), device, out.ctypes.data))
    return out


def triangulate_pts(Ps, Rs, tic, ric, start_frame, n_obs, obs_offset, obs, depth, init_depth: float = 5.0, device: int = -1) -> np.ndarray:
    """FeatureManager::TriangulatePts (feature_manager.cpp:202-275) -> estimated_depth_ per landmark (float64)."""
    w, keep = _window(Ps, Rs, tic, ric)
    sf, no, oo, ob = _table(start_frame, n_obs, obs_offset, obs)
    d = np.ascontiguousarray(depth, np.float64).copy()
    check(load().vf_op_triangulate_pts(C.byref(w), len(sf), sf.ctypes.data, no.ctypes.data, oo.ctypes.data, ob.ctypes.data, len(ob),
                                       float(init_depth), device, d.ctypes.data))
    return d


def outliers_rejection(Ps, Rs, tic, ric, start_frame, n_obs, obs_offset, obs, depth, focal_length: float = 460.0, device: int = -1):
    """FeatureManager::OutliersRejection (feature_manager.cpp:379-429) -> (mean reprojection error, remove flags)."""
    w, keep = _window(Ps, Rs, tic, ric)
    sf, no, oo, ob = _table(start_frame, n_obs, obs_offset, obs)
    d = np.ascontiguousarray(depth, np.float64)
    ave = np.zeros(len(sf), np.float64)
    rm = np.zeros(len(sf), np.uint8)
    check(load().vf_op_outliers_rejection(C.byref(w), len(sf), sf.ctypes.data, no.ctypes.data, oo.ctypes.data, ob.ctypes.data, len(ob),
                                          d.ctypes.data, float(focal_length), device, ave.ctypes.data, rm.ctypes.data))
    return ave, rm
