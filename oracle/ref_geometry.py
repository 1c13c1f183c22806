"""TEST INFRASTRUCTURE (oracle, kind "reference") -- ctypes binding of oracle/_ref/libvf_ref_geometry.so: the reference's own
FeatureManager::TriangulatePts / TriangulateOnePoint / ReProjectionError / OutliersRejection
(vins/estimator/feature_manager.cpp:202-293, 365-429), compiled from /root/reference by oracle/build_ref.py.
Only tests/ may import this; `available()` is False where neither the reference nor the prebuilt library exists."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import build_ref

_lib = None


def available() -> bool:
    return build_ref.build() is not None


def lib():
    global _lib
    if _lib is None:
        so = build_ref.build()
        if so is None:
            raise RuntimeError("oracle/_ref/libvf_ref_geometry.so is missing and /root/reference is not there to build it from")
        _lib = C.CDLL(so)
    return _lib


def _d(a, cols):
    return np.ascontiguousarray(a, np.float64).reshape(-1, cols)


def _i(a):
    return np.ascontiguousarray(a, np.int32)


def triangulate_one_point(pose0, pose1, point0, point1):
    a, b, p0, p1 = _d(pose0, 12), _d(pose1, 12), _d(point0, 2), _d(point1, 2)
    out = np.zeros((len(a), 3), np.float64)
    lib().vf_ref_triangulate_one_point(a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), p0.ctypes.data_as(C.c_void_p),
                                       p1.ctypes.data_as(C.c_void_p), C.c_int(len(a)), out.ctypes.data_as(C.c_void_p))
    return out


def _win(Ps, Rs, tic, ric):
    return _d(Ps, 3), _d(Rs, 9), _d(tic, 3), _d(ric, 9)


def triangulate_pts(Ps, Rs, tic, ric, start_frame, n_obs, obs_offset, obs, depth):
    Ps, Rs, tic, ric = _win(Ps, Rs, tic, ric)
    sf, no, oo, ob = _i(start_frame), _i(n_obs), _i(obs_offset), _d(obs, 7)
    d = np.ascontiguousarray(depth, np.float64).copy()
    p = lambda x: x.ctypes.data_as(C.c_void_p)
    lib().vf_ref_triangulate_pts(C.c_int(len(Ps)), p(Ps), p(Rs), p(tic), p(ric), C.c_int(len(sf)), p(sf), p(no), p(oo), p(ob), p(d))
    return d


def outliers_rejection(Ps, Rs, tic, ric, start_frame, n_obs, obs_offset, obs, depth, focal_length=460.0):
    Ps, Rs, tic, ric = _win(Ps, Rs, tic, ric)
    sf, no, oo, ob = _i(start_frame), _i(n_obs), _i(obs_offset), _d(obs, 7)
    d = np.ascontiguousarray(depth, np.float64)
    rm = np.zeros(len(sf), np.uint8)
    p = lambda x: x.ctypes.data_as(C.c_void_p)
    lib().vf_ref_outliers_rejection(C.c_int(len(Ps)), p(Ps), p(Rs), p(tic), p(ric), C.c_int(len(sf)), p(sf), p(no), p(oo), p(ob), p(d),
                                    C.c_double(focal_length), p(rm))
    return rm


def reprojection_error(Ri, Pi, rici, tici, Rj, Pj, ricj, ticj, depth, uvi, uvj):
    args = [_d(Ri, 9), _d(Pi, 3), _d(rici, 9), _d(tici, 3), _d(Rj, 9), _d(Pj, 3), _d(ricj, 9), _d(ticj, 3),
            np.ascontiguousarray(depth, np.float64), _d(uvi, 3), _d(uvj, 3)]
    n = len(args[0])
    err = np.zeros(n, np.float64)
    lib().vf_ref_reprojection_error(*[a.ctypes.data_as(C.c_void_p) for a in args], C.c_int(n), err.ctypes.data_as(C.c_void_p))
    return err
