"""Python mirror of the reference's estimator::FeatureTracker over the C ABI.

Same names and argument meaning as vins/estimator/feature_tracker.hpp:18-107:
  FeatureTracker(...)                 ctor reads max_cnt / min_dist / flow_back (feature_tracker.cpp:8-18)
  ReadIntrinsicParameter([cam0, cam1]) (feature_tracker.cpp:301-308)
  TrackImage(t, img0, img1) -> {feature_id: [(camera_id, [x, y, z, u, v, vx, vy]), ...]}  (feature_tracker.cpp:27-217)
The arithmetic runs on the GPU inside libvinsfast_b200.so; this file only marshals buffers.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _lib
from ._lib import FEATURE_DTYPE, VfCamera, VfConfig, VfError, check, load


def camera_struct(cam) -> VfCamera:
    """Accepts a VfCamera, a dict, or any object with model/xi/k1/k2/p1/p2/fx/fy/cx/cy attributes."""
    if isinstance(cam, VfCamera):
        return cam
    get = (lambda k, d=0.0: cam.get(k, d)) if isinstance(cam, dict) else (lambda k, d=0.0: getattr(cam, k, d))
    c = VfCamera()
    c.model = int(get("model", 0))
    c.image_width = int(get("image_width", 0))
    c.image_height = int(get("image_height", 0))
    for k in ("xi", "k1", "k2", "p1", "p2", "fx", "fy", "cx", "cy"):
        setattr(c, k, float(get(k, 0.0)))
    return c


def load_camera_yaml(path: str) -> VfCamera:
    c = VfCamera()
    check(load().vf_camera_load_yaml(path.encode(), C.byref(c)))
    return c


def default_config() -> VfConfig:
    cfg = VfConfig()
    load().vf_config_default(C.byref(cfg))
    return cfg


def load_config_yaml(path: str) -> VfConfig:
    cfg = VfConfig()
    check(load().vf_config_load_yaml(path.encode(), C.byref(cfg)))
    return cfg


def make_config(width: int, height: int, max_cnt: int, min_dist: int, flow_back: int = 1, cam0=None, cam1=None,
                lk_max_level: int = 3, quality_level: float = 0.01, device: int = -1, use_cuda_graph: bool = True,
                cuda_stream: int = 0) -> VfConfig:
    cfg = default_config()
    cfg.width, cfg.height, cfg.max_cnt, cfg.min_dist, cfg.flow_back = width, height, max_cnt, min_dist, flow_back
    cfg.lk_max_level, cfg.quality_level, cfg.device = lk_max_level, quality_level, device
    cfg.use_cuda_graph = 1 if use_cuda_graph else 0
    cfg.cuda_stream = cuda_stream or None
    if cam0 is not None:
        cfg.cam[0] = camera_struct(cam0)
    if cam1 is not None:
        cfg.cam[1] = camera_struct(cam1)
    return cfg


def _img(a) -> np.ndarray:
    a = np.asarray(a)
    if a.dtype != np.uint8 or a.ndim != 2:
        raise VfError(_lib.VF_ERR_INVALID_ARG, "images must be 2-D uint8 (CV_8UC1)")
    if a.strides[1] != 1:
        a = np.ascontiguousarray(a)
    return a


def feature_frame(rec: np.ndarray) -> dict:
    """vf_feature[] -> the reference's featureFrame map (feature_tracker.cpp:173-213): left observations in ids_ order,
    then right observations appended for the ids that have one."""
    ff: dict = {}
    for r in rec:
        ff.setdefault(int(r["id"]), []).append(
            (0, np.array([r["x"], r["y"], 1.0, r["u"], r["v"], r["vx"], r["vy"]], np.float64)))
    for r in rec[rec["has_right"] != 0]:
        ff.setdefault(int(r["id"]), []).append(
            (1, np.array([r["xr"], r["yr"], 1.0, r["ur"], r["vr"], r["vxr"], r["vyr"]], np.float64)))
    return ff


class FeatureTracker:
    """One stream of stereo frames on one GPU."""

    def __init__(self, config: Optional[VfConfig] = None, *, width: int = 752, height: int = 480, max_cnt: int = 200,
                 min_dist: int = 15, flow_back: int = 1, lk_max_level: int = 3, quality_level: float = 0.01,
                 device: int = -1, use_cuda_graph: bool = True, cuda_stream: int = 0):
        self._lib = load()
        self._h = C.c_void_p()
        self.cfg = config if config is not None else make_config(
            width, height, max_cnt, min_dist, flow_back, None, None, lk_max_level, quality_level, device,
            use_cuda_graph, cuda_stream)
        # public fields of the reference class (feature_tracker.hpp:84-90)
        self.MAX_CNT_FEATURES_PER_FRAME = self.cfg.max_cnt
        self.MIN_DIST = self.cfg.min_dist
        self.OPTICAL_FLOW_BACK = self.cfg.flow_back
        self.SHOW_TRACK_RESULT = 0
        self._out = np.zeros(max(self.cfg.max_cnt, 1), FEATURE_DTYPE)
        self._n = C.c_int32(0)

    # -- reference API ---------------------------------------------------------------------------------
    def ReadIntrinsicParameter(self, calib_file_path: Sequence):
        """Index 0 = left, 1 = right; entries are YAML paths or camera descriptions."""
        for k, c in enumerate(list(calib_file_path)[:2]):
            self.cfg.cam[k] = load_camera_yaml(c) if isinstance(c, str) else camera_struct(c)
        self._destroy()

    def TrackImage(self, current_time: float, img0, img1) -> dict:
        return feature_frame(self.track(current_time, img0, img1))

    # -- array API -------------------------------------------------------------------------------------
    def _ensure(self):
        if not self._h:
            h = C.c_void_p()
            check(self._lib.vf_tracker_create(C.byref(self.cfg), C.byref(h)))
            self._h = h

    def track(self, t: float, img0, img1) -> np.ndarray:
        """Returns the vf_feature records (structured array, ids_ order)."""
        if img0 is None or img1 is None:
            raise VfError(_lib.VF_ERR_INVALID_ARG, "img0/img1 must both be given (stereo is mandatory, feature_tracker.cpp:30)")
        a, b = _img(img0), _img(img1)
        if a.shape != (self.cfg.height, self.cfg.width) or b.shape != a.shape:
            raise VfError(_lib.VF_ERR_INVALID_ARG, f"image shape {a.shape}/{b.shape} != configured {(self.cfg.height, self.cfg.width)}")
        self._ensure()
        st = self._lib.vf_tracker_track(self._h, float(t), a.ctypes.data, a.strides[0], b.ctypes.data, b.strides[0],
                                        self._out.ctypes.data, C.byref(self._n))
        check(st, self.last_error)
        return self._out[: self._n.value].copy()

    def track_device(self, t: float, d_img0: int, d_img1: int, row_stride: int) -> np.ndarray:
        """Images already on the device (raw device pointers, e.g. torch_tensor.data_ptr())."""
        self._ensure()
        st = self._lib.vf_tracker_track_device(self._h, float(t), d_img0, row_stride, d_img1, row_stride,
                                               self._out.ctypes.data, C.byref(self._n))
        check(st, self.last_error)
        return self._out[: self._n.value].copy()

    def reset(self):
        if self._h:
            check(self._lib.vf_tracker_reset(self._h), self.last_error)

    def last_error(self) -> str:
        return (self._lib.vf_last_error(self._h) or b"").decode() if self._h else _lib.global_error()

    def debug(self, what: int, arg: int = 0, dtype=np.uint8, shape=None) -> np.ndarray:
        return _debug_get(self._lib, self._h, what, arg, dtype, shape, self.cfg)

    def _destroy(self):
        if getattr(self, "_h", None):
            self._lib.vf_tracker_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self._destroy()
        except Exception:
            pass


def _debug_get(lib, handle, what, arg, dtype, shape, cfg) -> np.ndarray:
    size = C.c_size_t(0)
    cap = max(cfg.width * cfg.height * 4, cfg.max_cnt * 16, 64)
    buf = np.zeros(cap, np.uint8)
    st = lib.vf_tracker_debug_get(handle, what, arg, buf.ctypes.data, buf.nbytes, C.byref(size))
    check(st, lambda: (lib.vf_last_error(handle) or b"").decode())
    out = buf[: size.value].view(dtype).copy()
    return out.reshape(shape) if shape is not None else out


class Batch:
    """S independent stereo streams advanced in lock-step on one GPU (BASELINE config 4)."""

    def __init__(self, config: VfConfig, n_streams: int):
        self._lib = load()
        self.cfg, self.n_streams = config, n_streams
        self._h = C.c_void_p()
        check(self._lib.vf_batch_create(C.byref(self.cfg), n_streams, C.byref(self._h)))
        self._out = np.zeros((n_streams, max(config.max_cnt, 1)), FEATURE_DTYPE)
        self._n = np.zeros(n_streams, np.int32)
        self._tbuf = np.zeros(n_streams, np.float64)           # per-call scratch of the fast path (addresses taken once)
        self._tbuf_p, self._out_p, self._n_p = self._tbuf.ctypes.data, self._out.ctypes.data, self._n.ctypes.data

    def last_error(self) -> str:
        return (self._lib.vf_batch_last_error(self._h) or b"").decode()

    def _times(self, times):
        t = np.ascontiguousarray(np.broadcast_to(np.asarray(times, np.float64), (self.n_streams,)))
        return t

    def _results(self):
        return [self._out[s, : self._n[s]].copy() for s in range(self.n_streams)]

    def track(self, times, imgs0, imgs1):
        a = [_img(x) for x in imgs0]
        b = [_img(x) for x in imgs1]
        assert len(a) == len(b) == self.n_streams
        stride0, stride1 = a[0].strides[0], b[0].strides[0]
        assert all(x.strides[0] == stride0 for x in a) and all(x.strides[0] == stride1 for x in b)
        pa = (C.c_void_p * self.n_streams)(*[x.ctypes.data for x in a])
        pb = (C.c_void_p * self.n_streams)(*[x.ctypes.data for x in b])
        t = self._times(times)
        st = self._lib.vf_batch_track(self._h, t.ctypes.data, pa, stride0, pb, stride1, self._out.ctypes.data,
                                      self._n.ctypes.data)
        check(st, self.last_error)
        return self._results()

    def track_ptrs(self, times, ptrs0, stride0: int, ptrs1, stride1: int):
        """Host pointer arrays prepared once by the caller (ctypes arrays of c_void_p): the per-frame fast path."""
        self._tbuf[:] = times
        st = self._lib.vf_batch_track(self._h, self._tbuf_p, ptrs0, stride0, ptrs1, stride1, self._out_p, self._n_p)
        if st:
            check(st, self.last_error)
        return self._n

    def enqueue_device(self, times, d_img0: int, d_img1: int, row_stride: int, image_stride: int):
        t = self._times(times)
        check(self._lib.vf_batch_enqueue_device(self._h, t.ctypes.data, d_img0, d_img1, row_stride, image_stride),
              self.last_error)

    def fetch(self, lag: int = 0):
        """Results of the last enqueued frame (lag 0) or of the ones before it (lag 1 or 2, while the later ones still run)."""
        check(self._lib.vf_batch_fetch_lagged(self._h, lag, self._out.ctypes.data, self._n.ctypes.data), self.last_error)
        return self._results()

    def enqueue_ptrs(self, times, ptrs0, stride0: int, ptrs1, stride1: int):
        """Starts a frame from HOST images (ctypes arrays of c_void_p) and returns at once; see vf_batch_enqueue."""
        t = self._times(times)
        check(self._lib.vf_batch_enqueue(self._h, t.ctypes.data, ptrs0, stride0, ptrs1, stride1), self.last_error)

    def fetch_counts_lagged(self, lag: int) -> np.ndarray:
        check(self._lib.vf_batch_fetch_lagged(self._h, lag, self._out.ctypes.data, self._n.ctypes.data), self.last_error)
        return self._n

    def fetch_counts(self) -> np.ndarray:
        check(self._lib.vf_batch_fetch(self._h, self._out.ctypes.data, self._n.ctypes.data), self.last_error)
        return self._n

    def track_device(self, times, d_img0: int, d_img1: int, row_stride: int, image_stride: int):
        self.enqueue_device(times, d_img0, d_img1, row_stride, image_stride)
        return self.fetch()

    def reset(self):
        check(self._lib.vf_batch_reset(self._h), self.last_error)

    @property
    def cuda_stream(self) -> int:
        return int(self._lib.vf_batch_cuda_stream(self._h) or 0)

    def host_times_us(self):
        """Cumulative host microseconds inside the enqueue calls: (waiting for frame k-4, issuing the frames)."""
        t = (C.c_double * 2)()
        self._lib.vf_batch_host_times(self._h, t)
        return float(t[0]), float(t[1])

    @property
    def launches_per_frame(self) -> int:
        return int(self._lib.vf_batch_launches_per_frame(self._h))

    def set_profiling(self, enable: bool):
        check(self._lib.vf_batch_set_profiling(self._h, 1 if enable else 0), self.last_error)

    def stage_times_ms(self) -> dict:
        n = self._lib.vf_batch_stage_count()
        ms = np.zeros(n, np.float32)
        check(self._lib.vf_batch_stage_times(self._h, ms.ctypes.data, n), self.last_error)
        return {self._lib.vf_batch_stage_name(i).decode(): float(ms[i]) for i in range(n)}

    def debug(self, stream: int, what: int, arg: int = 0, dtype=np.uint8, shape=None) -> np.ndarray:
        view = self._lib.vf_batch_stream(self._h, stream)
        return _debug_get(self._lib, view, what, arg, dtype, shape, self.cfg)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.vf_batch_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
