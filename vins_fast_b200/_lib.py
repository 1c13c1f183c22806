"""ctypes binding of include/vins_fast_b200.h (the C ABI of libvinsfast_b200.so).

The library is built in-tree by vins_fast_b200/build.py.  There is no CPU fallback: if the shared
object is missing this module raises, and every compute call fails loudly without a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VF_LIB") or os.path.join(HERE, "libvinsfast_b200.so")   # VF_LIB: a measurement build of the same library

VF_OK, VF_ERR_INVALID_ARG, VF_ERR_CUDA, VF_ERR_IO, VF_ERR_PARSE, VF_ERR_STATE = range(6)
VF_CAM_MEI, VF_CAM_PINHOLE = 0, 1
STATUS_NAMES = {0: "VF_OK", 1: "VF_ERR_INVALID_ARG", 2: "VF_ERR_CUDA", 3: "VF_ERR_IO", 4: "VF_ERR_PARSE", 5: "VF_ERR_STATE"}

(DBG_PYR_LEFT, DBG_PYR_RIGHT, DBG_EIG, DBG_MASK, DBG_FWD_PTS, DBG_FWD_STATUS, DBG_REV_PTS, DBG_REV_STATUS,
 DBG_TEMPORAL_STATUS, DBG_SORT_PERM, DBG_LR_PTS, DBG_LR_STATUS, DBG_RL_PTS, DBG_RL_STATUS, DBG_COUNTS,
 DBG_LK_ITERS) = range(16)


class VfCamera(C.Structure):
    _fields_ = [("model", C.c_int32), ("image_width", C.c_int32), ("image_height", C.c_int32)] + \
               [(n, C.c_double) for n in ("xi", "k1", "k2", "p1", "p2", "fx", "fy", "cx", "cy")]


class VfConfig(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("max_cnt", C.c_int32), ("min_dist", C.c_int32),
                ("flow_back", C.c_int32), ("lk_max_level", C.c_int32), ("lk_win", C.c_int32),
                ("quality_level", C.c_double), ("cam", VfCamera * 2), ("device", C.c_int32),
                ("use_cuda_graph", C.c_int32), ("cuda_stream", C.c_void_p)]


class VfWindow(C.Structure):
    _fields_ = [("window", C.c_int32), ("Ps", C.c_void_p), ("Rs", C.c_void_p), ("tic", (C.c_double * 3) * 2), ("ric", (C.c_double * 9) * 2)]


class VfFeature(C.Structure):
    _fields_ = [("id", C.c_int32), ("track_cnt", C.c_int32), ("has_right", C.c_int32)] + \
               [(n, C.c_float) for n in ("x", "y", "u", "v", "vx", "vy", "xr", "yr", "ur", "vr", "vxr", "vyr")]


FEATURE_DTYPE = np.dtype([("id", "<i4"), ("track_cnt", "<i4"), ("has_right", "<i4")] +
                         [(n, "<f4") for n in ("x", "y", "u", "v", "vx", "vy", "xr", "yr", "ur", "vr", "vxr", "vyr")])
assert FEATURE_DTYPE.itemsize == C.sizeof(VfFeature) == 60

# every symbol include/vins_fast_b200.h declares (tests check that the library exports all of them)
EXPORTS = [
    "vf_config_default", "vf_config_load_yaml", "vf_camera_load_yaml", "vf_yaml_get", "vf_last_global_error",
    "vf_tracker_create", "vf_tracker_destroy", "vf_tracker_reset", "vf_tracker_track", "vf_tracker_track_device",
    "vf_last_error",
    "vf_batch_create", "vf_batch_destroy", "vf_batch_reset", "vf_batch_size", "vf_batch_track", "vf_batch_track_device",
    "vf_batch_enqueue_device", "vf_batch_set_input_event", "vf_batch_fetch", "vf_batch_enqueue", "vf_batch_fetch_lagged", "vf_batch_last_error", "vf_batch_cuda_stream",
    "vf_batch_launches_per_frame", "vf_batch_host_times", "vf_batch_stream",
    "vf_batch_set_profiling", "vf_batch_stage_count", "vf_batch_stage_name", "vf_batch_stage_times", "vf_batch_trace_get",
    "vf_tracker_debug_get",
    "vf_debug_set_response_kernel", "vf_debug_set_pyramid_kernel", "vf_debug_set_lk_kernel",
    "vf_op_build_pyramid", "vf_op_scharr", "vf_op_lk", "vf_op_lk_win", "vf_op_min_eigen", "vf_op_good_features", "vf_op_set_mask",
    "vf_op_undistort", "vf_op_triangulate_one_point", "vf_op_triangulate_pts", "vf_op_outliers_rejection",
    "vf_host_alloc", "vf_host_free", "vf_device_count",
]

_lib = None


class VfError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"{STATUS_NAMES.get(status, status)}: {message}")
        self.status = status
        self.message = message


def load() -> C.CDLL:
    """Loads the in-tree CUDA library; never substitutes anything else."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python -m vins_fast_b200.build` "
                          "(there is no CPU fallback for the tracker)")
    lib = C.CDLL(LIB_PATH)
    P, I32, F64, SZ = C.c_void_p, C.c_int32, C.c_double, C.c_size_t
    sig = {
        "vf_config_default": (None, [P]),
        "vf_config_load_yaml": (I32, [C.c_char_p, P]),
        "vf_camera_load_yaml": (I32, [C.c_char_p, P]),
        "vf_yaml_get": (I32, [C.c_char_p, C.c_char_p, P, SZ]),
        "vf_last_global_error": (C.c_char_p, []),
        "vf_tracker_create": (I32, [P, P]),
        "vf_tracker_destroy": (None, [P]),
        "vf_tracker_reset": (I32, [P]),
        "vf_tracker_track": (I32, [P, F64, P, SZ, P, SZ, P, P]),
        "vf_tracker_track_device": (I32, [P, F64, P, SZ, P, SZ, P, P]),
        "vf_last_error": (C.c_char_p, [P]),
        "vf_batch_create": (I32, [P, I32, P]),
        "vf_batch_destroy": (None, [P]),
        "vf_batch_reset": (I32, [P]),
        "vf_batch_size": (I32, [P]),
        "vf_batch_track": (I32, [P, P, P, SZ, P, SZ, P, P]),
        "vf_batch_track_device": (I32, [P, P, P, P, SZ, SZ, P, P]),
        "vf_batch_enqueue_device": (I32, [P, P, P, P, SZ, SZ]),
        "vf_batch_fetch": (I32, [P, P, P]),
        "vf_batch_set_input_event": (I32, [P, P]),
        "vf_batch_enqueue": (I32, [P, P, P, SZ, P, SZ]),
        "vf_batch_fetch_lagged": (I32, [P, I32, P, P]),
        "vf_batch_last_error": (C.c_char_p, [P]),
        "vf_batch_cuda_stream": (P, [P]),
        "vf_batch_launches_per_frame": (I32, [P]),
        "vf_batch_host_times": (None, [P, P]),
        "vf_batch_stream": (P, [P, I32]),
        "vf_batch_set_profiling": (I32, [P, I32]),
        "vf_batch_stage_count": (I32, []),
        "vf_batch_stage_name": (C.c_char_p, [I32]),
        "vf_batch_stage_times": (I32, [P, P, I32]),
        "vf_batch_trace_get": (I32, [P, P, I32]),
        "vf_tracker_debug_get": (I32, [P, I32, I32, P, SZ, P]),
        "vf_op_build_pyramid": (I32, [P, I32, I32, I32, I32, P, P, P, P]),
        "vf_op_scharr": (I32, [P, I32, I32, I32, P]),
        "vf_op_lk": (I32, [P, P, I32, I32, P, P, P, I32, I32, I32, I32, P]),
        "vf_op_lk_win": (I32, [P, P, I32, I32, P, P, P, I32, I32, I32, I32, I32, P]),
        "vf_debug_set_response_kernel": (I32, [I32]),
        "vf_debug_set_pyramid_kernel": (I32, [I32]),
        "vf_debug_set_lk_kernel": (I32, [I32]),
        "vf_op_min_eigen": (I32, [P, I32, I32, I32, P]),
        "vf_op_good_features": (I32, [P, I32, I32, I32, F64, I32, P, I32, P, P]),
        "vf_op_set_mask": (I32, [P, P, I32, I32, I32, I32, I32, P, P, P, P]),
        "vf_op_undistort": (I32, [P, P, I32, I32, P]),
        "vf_op_triangulate_one_point": (I32, [P, P, P, P, I32, I32, P]),
        "vf_op_triangulate_pts": (I32, [P, I32, P, P, P, P, I32, F64, I32, P]),
        "vf_op_outliers_rejection": (I32, [P, I32, P, P, P, P, I32, P, F64, I32, P, P]),
        "vf_host_alloc": (I32, [SZ, P]),
        "vf_host_free": (None, [P]),
        "vf_device_count": (I32, []),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def global_error() -> str:
    return (load().vf_last_global_error() or b"").decode()


def check(status: int, message_fn=None):
    if status != VF_OK:
        msg = message_fn() if message_fn else global_error()
        raise VfError(status, msg)
