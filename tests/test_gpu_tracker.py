"""Whole-path parity: vf_tracker_track (C ABI) against the CPU oracle on identical synthetic stereo sequences.
The oracle's TrackImage (oracle/vf_oracle.c) is bit-identical to the cv2-based restatement of the reference
(tests/test_oracle_golden.py), and the GPU path must be bit-identical to the oracle: ids, track counts, pixel
coordinates, undistorted coordinates and velocities, frame after frame (north_star tolerances -- 0.01 px,
99.5 % status agreement, 1e-6 undistorted -- are therefore met with zero slack)."""
import ctypes as C

import numpy as np
import pytest

from conftest import assert_frames_equal
from oracle.camera import EUROC_CAM0, EUROC_CAM1, pinhole_for

pytestmark = pytest.mark.gpu


def _tracker(w, h, max_cnt, min_dist, flow_back, cam0, cam1, level=3, graph=True, win=21):
    import vins_fast_b200 as v
    cfg = v.make_config(w, h, max_cnt, min_dist, flow_back, cam0.__dict__, cam1.__dict__, lk_max_level=level,
                        use_cuda_graph=graph)
    cfg.lk_win = win
    return v.FeatureTracker(cfg)


def _run_pair(oracle, frames, w, h, max_cnt, min_dist, flow_back, cam0, cam1, level=3, graph=True, check_debug=False, win=21):
    from vins_fast_b200 import _lib
    gpu = _tracker(w, h, max_cnt, min_dist, flow_back, cam0, cam1, level, graph, win)
    ref = oracle.CFeatureTracker(cam0, cam1, w, h, max_cnt, min_dist, flow_back, level, win, 0.01, 0)
    total = 0
    for k, (t, l, r) in enumerate(frames):
        g = oracle.records_to_frame(gpu.track(t, l, r))
        o = ref.track_image(t, l, r)
        assert_frames_equal(g, o, f"frame {k}")
        total += len(g["ids"])
        if check_debug:
            cnts = gpu.debug(_lib.DBG_COUNTS, dtype=np.int32)
            assert cnts[4] == len(g["ids"]) and cnts[5] == len(g["ids_right"])
            for lvl in range(4):
                assert np.array_equal(gpu.debug(_lib.DBG_PYR_LEFT, lvl), oracle.Pyramid(l).level(lvl).ravel())
                assert np.array_equal(gpu.debug(_lib.DBG_PYR_RIGHT, lvl), oracle.Pyramid(r).level(lvl).ravel())
            eig = gpu.debug(_lib.DBG_EIG, dtype=np.float32).reshape(h, w)
            assert np.array_equal(eig.view(np.uint32), oracle.min_eigen_val(l).view(np.uint32))
    return total


def test_euroc_sequence_bit_exact(lib, oracle, euroc_frames):
    n = _run_pair(oracle, euroc_frames, 752, 480, 150, 30, 1, EUROC_CAM0, EUROC_CAM1)
    assert n > 20 * 140


def test_euroc_reference_defaults_200_15(lib, oracle, euroc_frames):
    _run_pair(oracle, euroc_frames[:12], 752, 480, 200, 15, 1, EUROC_CAM0, EUROC_CAM1)


def test_euroc_no_graph_and_debug_getters(lib, oracle, euroc_frames):
    _run_pair(oracle, euroc_frames[:4], 752, 480, 150, 30, 1, EUROC_CAM0, EUROC_CAM1, graph=False, check_debug=True)


def test_flow_back_off(lib, oracle, euroc_frames):
    _run_pair(oracle, euroc_frames[:8], 752, 480, 150, 30, 0, EUROC_CAM0, EUROC_CAM1)


@pytest.mark.parametrize("win,level,flow_back", [(15, 3, 1), (31, 3, 1), (9, 2, 0), (27, 4, 1)])
def test_euroc_sequence_with_other_lk_windows(lib, oracle, euroc_frames, win, level, flow_back):
    """`lk_win` is a parameter (north star); the reference's literal is 21.  Every other odd width 5..31 runs the
    run-time-width kernels of csrc/vf_lk_generic.cu: the whole TrackImage stays bit-identical to the oracle with that window
    (and the oracle to cv2 with that window: tests/test_oracle_golden.py)."""
    n = _run_pair(oracle, euroc_frames[:10], 752, 480, 150, 30, flow_back, EUROC_CAM0, EUROC_CAM1, level=level, win=win)
    assert n > 10 * 120


def test_other_lk_window_batch_and_golden_sequences(lib, oracle):
    """Window 15 / 31 against the committed cv2 fixture directly (GPU vs OpenCV, no oracle in between), through a batch of two
    streams so that the stream index of the run-time-width kernels is exercised as well."""
    import os
    import vins_fast_b200 as v
    from vins_fast_b200.synth import StereoSequence
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "lk_windows_cv2_4_13.npz"))
    for name in ("win15", "win31"):
        w, h, max_cnt, min_dist, flow_back, level, n_frames, seed, win = (int(x) for x in g[f"{name}_params"])
        cam = pinhole_for(w, h)
        cfg = v.make_config(w, h, max_cnt, min_dist, flow_back, cam.__dict__, cam.__dict__, lk_max_level=level)
        cfg.lk_win = win
        batch = v.Batch(cfg, 2)
        seq = StereoSequence(seed, w, h)
        for k in range(n_frames):
            t, l, r = seq.frame(k)
            recs = batch.track(t, np.stack([l, l]), np.stack([r, r]))
            for s in range(2):
                f = oracle.records_to_frame(recs[s])
                for key, val in f.items():
                    ref = g[f"{name}_f{k}_{key}"]
                    assert val.shape == ref.shape and val.tobytes() == ref.astype(val.dtype).tobytes(), (name, k, s, key)


@pytest.mark.parametrize("mode", [1, 2], ids=["fused", "tma"])
def test_euroc_sequence_with_either_pyramid_kernel(lib, oracle, euroc_frames, mode):
    """The whole frame (level 0 moved into its padded home by the pyramid launch, borders, both cameras) with the pyramid
    kernel forced: 752x480 normally takes the fused kernel, 4K and batches the streaming TMA kernel."""
    from vins_fast_b200 import ops
    ops.set_pyramid_kernel(mode)
    try:
        _run_pair(oracle, euroc_frames[:6], 752, 480, 150, 30, 1, EUROC_CAM0, EUROC_CAM1, check_debug=(mode == 2))
        frames = [(t, np.ascontiguousarray(l[:301, :433]), np.ascontiguousarray(r[:301, :433])) for t, l, r in euroc_frames[:4]]
        cam = pinhole_for(433, 301, distorted=False)
        _run_pair(oracle, frames, 433, 301, 40, 25, 1, cam, cam)
    finally:
        ops.set_pyramid_kernel(0)


@pytest.mark.parametrize("mode", [1, 2, 3], ids=["cta", "warp", "coop"])
def test_euroc_sequence_with_either_lk_kernel(lib, oracle, euroc_frames, mode):
    """The tracker with the LK kernels forced: block-barrier kernels of vf_lk.cu / warp-level passes with one warp per feature
    (what batches take) / with one cooperating CTA per feature (what a single stream takes)."""
    from vins_fast_b200 import ops
    ops.set_lk_kernel(mode)
    try:
        _run_pair(oracle, euroc_frames[:8], 752, 480, 150, 30, 1, EUROC_CAM0, EUROC_CAM1)
    finally:
        ops.set_lk_kernel(0)


def test_pinhole_720p_level4(lib, oracle):
    from vins_fast_b200.synth import StereoSequence
    seq = StereoSequence(2, 1280, 720)
    cam = pinhole_for(1280, 720)
    _run_pair(oracle, [seq.frame(k) for k in range(5)], 1280, 720, 400, 30, 1, cam, cam, level=4)


def test_odd_size_and_small_max_cnt(lib, oracle, euroc_seq):
    frames = [(t, np.ascontiguousarray(l[:301, :433]), np.ascontiguousarray(r[:301, :433]))
              for t, l, r in (euroc_seq.frame(k) for k in range(6))]
    cam = pinhole_for(433, 301, distorted=False)
    _run_pair(oracle, frames, 433, 301, 20, 25, 1, cam, cam)


def test_featureframe_contract_and_strides(lib, oracle, euroc_frames):
    """featureFrame invariants the consumer asserts (feature_manager.cpp:36-43) + strided (non-contiguous) input."""
    import vins_fast_b200 as v
    tr = _tracker(752, 480, 150, 30, 1, EUROC_CAM0, EUROC_CAM1)
    pad = np.zeros((480, 800), np.uint8)
    ref = oracle.CFeatureTracker(EUROC_CAM0, EUROC_CAM1, 752, 480, 150, 30, 1)
    for t, l, r in euroc_frames[:3]:
        pl, pr = pad.copy(), pad.copy()
        pl[:, :752], pr[:, :752] = l, r
        ff = tr.TrackImage(t, pl[:, :752], pr[:, :752])          # row stride 800
        o = ref.track_image(t, l, r)
        assert sorted(ff.keys()) == sorted(int(i) for i in o["ids"])
        for fid, obs in ff.items():
            assert 1 <= len(obs) <= 2 and obs[0][0] == 0 and obs[0][1][2] == 1.0
            if len(obs) == 2:
                assert obs[1][0] == 1 and obs[1][1][2] == 1.0
        assert sum(len(o_) == 2 for o_ in ff.values()) == len(o["ids_right"])


def test_reset_restarts_ids(lib, oracle, euroc_frames):
    tr = _tracker(752, 480, 150, 30, 1, EUROC_CAM0, EUROC_CAM1)
    a = [tr.track(t, l, r) for t, l, r in euroc_frames[:3]]
    tr.reset()
    b = [tr.track(t, l, r) for t, l, r in euroc_frames[:3]]
    for x, y in zip(a, b):
        assert x.tobytes() == y.tobytes()
    assert a[0]["id"].min() == 0 and np.all(a[0]["vx"] == 0)


def test_batch_streams_independent_and_deterministic(lib, oracle):
    """BASELINE config 4 in small: S streams in one batch == S single-stream runs == oracle."""
    import vins_fast_b200 as v
    from vins_fast_b200.synth import StereoSequence
    S, nfr = 3, 5
    seqs = [StereoSequence(10 + s, 752, 480) for s in range(S)]
    cfg = v.make_config(752, 480, 150, 30, 1, EUROC_CAM0.__dict__, EUROC_CAM1.__dict__)
    batch = v.Batch(cfg, S)
    refs = [oracle.CFeatureTracker(EUROC_CAM0, EUROC_CAM1, 752, 480, 150, 30, 1) for _ in range(S)]
    for k in range(nfr):
        fr = [q.frame(k) for q in seqs]
        res = batch.track([f[0] for f in fr], [f[1] for f in fr], [f[2] for f in fr])
        for s in range(S):
            assert_frames_equal(oracle.records_to_frame(res[s]), refs[s].track_image(*fr[s]), f"stream {s} frame {k}")
    assert batch.launches_per_frame >= 10
    batch.close()


def test_device_resident_input(lib, oracle, euroc_frames):
    torch = pytest.importorskip("torch")
    tr = _tracker(752, 480, 150, 30, 1, EUROC_CAM0, EUROC_CAM1)
    ref = oracle.CFeatureTracker(EUROC_CAM0, EUROC_CAM1, 752, 480, 150, 30, 1)
    for t, l, r in euroc_frames[:4]:
        dl, dr = torch.from_numpy(l).cuda(), torch.from_numpy(r).cuda()
        torch.cuda.synchronize()
        g = oracle.records_to_frame(tr.track_device(t, dl.data_ptr(), dr.data_ptr(), 752))
        assert_frames_equal(g, ref.track_image(t, l, r))


def test_error_codes(lib):
    import vins_fast_b200 as v
    from vins_fast_b200 import _lib
    cfg = v.make_config(752, 480, 150, 30)
    cfg.lk_win = 16                                                        # even / outside [5, 31]: no such OpenCV window here
    with pytest.raises(v.VfError) as e:
        v.FeatureTracker(cfg)._ensure()
    assert e.value.status == _lib.VF_ERR_INVALID_ARG
    cfg = v.make_config(32, 32, 150, 30)
    with pytest.raises(v.VfError):
        v.FeatureTracker(cfg)._ensure()
    tr = v.FeatureTracker(v.make_config(752, 480, 150, 30))
    with pytest.raises(v.VfError):
        tr.track(0.0, np.zeros((480, 752), np.uint8), None)              # stereo is mandatory (feature_tracker.cpp:30)
    with pytest.raises(v.VfError):
        tr.track(0.0, np.zeros((100, 100), np.uint8), np.zeros((100, 100), np.uint8))
    tr._ensure()
    out = np.zeros(150, v.FEATURE_DTYPE)
    n = C.c_int32(0)
    img = np.zeros((480, 752), np.uint8)
    st = lib.vf_tracker_track(tr._h, 0.0, img.ctypes.data, 100, img.ctypes.data, 752, out.ctypes.data, C.byref(n))
    assert st == _lib.VF_ERR_INVALID_ARG and b"stride" in lib.vf_last_error(tr._h)
    st = lib.vf_tracker_track(tr._h, 0.0, None, 752, img.ctypes.data, 752, out.ctypes.data, C.byref(n))
    assert st == _lib.VF_ERR_INVALID_ARG
    size = C.c_size_t(0)
    st = lib.vf_tracker_debug_get(tr._h, _lib.DBG_EIG, 0, None, 0, C.byref(size))
    assert st == _lib.VF_ERR_STATE                                        # no frame yet


def test_cpp_feature_tracker_drop_in(lib, oracle, euroc_frames, tmp_path):
    """The header-only C++ class (include/vins_fast_b200/feature_tracker.hpp), driven like Estimator::FrontendTracker
    (estimator.cpp:143-158) from YAML files: its featureFrame maps equal the oracle's, value for value."""
    import struct
    import subprocess
    from conftest import build_cpp_test, write_camera_yaml, write_estimator_yaml
    exe = build_cpp_test("test_feature_tracker")
    write_camera_yaml(tmp_path / "cam0_mei.yaml", EUROC_CAM0, 752, 480)
    write_camera_yaml(tmp_path / "cam1_mei.yaml", EUROC_CAM1, 752, 480)
    est = write_estimator_yaml(tmp_path / "est.yaml", 150, 30, 1, show_track=1)
    n = 6
    with open(tmp_path / "frames.bin", "wb") as f:
        for t, l, r in euroc_frames[:n]:
            f.write(struct.pack("<d", t)); f.write(l.tobytes()); f.write(r.tobytes())
    res = subprocess.run([exe, "run", str(est), str(tmp_path / "cam0_mei.yaml"), str(tmp_path / "cam1_mei.yaml"),
                          str(tmp_path / "frames.bin"), "752", "480", str(n), str(tmp_path / "out.bin")],
                         capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    buf = open(tmp_path / "out.bin", "rb").read()
    off = 0
    from oracle.cv2_oracle import Cv2FeatureTracker
    ref = oracle.CFeatureTracker(EUROC_CAM0, EUROC_CAM1, 752, 480, 150, 30, 1)
    for k, (t, l, r) in enumerate(euroc_frames[:n]):
        want = Cv2FeatureTracker.feature_frame(ref.track_image(t, l, r))
        (n_ids,) = struct.unpack_from("<i", buf, off); off += 4
        assert n_ids == len(want)
        for fid in sorted(want):                                   # std::map iterates in key order
            gid, n_obs = struct.unpack_from("<ii", buf, off); off += 8
            assert gid == fid and n_obs == len(want[fid])
            for cam_id, vec in want[fid]:
                (gcam,) = struct.unpack_from("<i", buf, off); off += 4
                v = np.frombuffer(buf, np.float64, 7, off); off += 56
                assert gcam == cam_id and v.tobytes() == vec.tobytes(), f"frame {k} id {fid} cam {cam_id}"
    assert off == len(buf)


def test_pipelined_enqueue_fetch_lagged(lib, oracle, euroc_frames):
    """vf_batch_enqueue + vf_batch_fetch_lagged(1): frame k-1 is collected while frame k runs; results are the same
    records the synchronous call returns (= the oracle's), frame after frame."""
    import vins_fast_b200 as v
    from vins_fast_b200 import _lib
    cfg = v.make_config(752, 480, 150, 30, 1, EUROC_CAM0.__dict__, EUROC_CAM1.__dict__)
    batch = v.Batch(cfg, 1)
    ref = oracle.CFeatureTracker(EUROC_CAM0, EUROC_CAM1, 752, 480, 150, 30, 1)
    with pytest.raises(v.VfError) as e:
        batch.fetch(1)
    assert e.value.status == _lib.VF_ERR_STATE
    want = []
    frames = euroc_frames[:10]
    keep = []
    for k, (t, l, r) in enumerate(frames):
        want.append(ref.track_image(t, l, r))
        pl, pr = (C.c_void_p * 1)(l.ctypes.data), (C.c_void_p * 1)(r.ctypes.data)
        keep.append((pl, pr))
        batch.enqueue_ptrs(t, pl, 752, pr, 752)
        if k >= 1:
            assert_frames_equal(oracle.records_to_frame(batch.fetch(1)[0]), want[k - 1], f"lagged frame {k - 1}")
    assert_frames_equal(oracle.records_to_frame(batch.fetch(0)[0]), want[-1], "last frame")
    # lag 2: two frames stay in flight behind the one being collected
    batch.reset()
    with pytest.raises(v.VfError):
        batch.fetch(3)
    for k, (t, l, r) in enumerate(frames):
        batch.enqueue_ptrs(t, *keep[k][:1], 752, *keep[k][1:], 752)
        if k >= 2:
            assert_frames_equal(oracle.records_to_frame(batch.fetch(2)[0]), want[k - 2], f"lag-2 frame {k - 2}")
    assert_frames_equal(oracle.records_to_frame(batch.fetch(1)[0]), want[-2], "lag-2 run, frame before last")
    assert_frames_equal(oracle.records_to_frame(batch.fetch(0)[0]), want[-1], "lag-2 run, last frame")
    batch.close()


@pytest.mark.parametrize("win", [15, 31])
def test_other_lk_window_overlapped_plan(lib, oracle, euroc_frames, win):
    """The run-time-width LK kernels under the OVERLAPPED frame plan (vf_batch_enqueue + vf_batch_fetch_lagged: the stereo pass
    of frame k beside the chain of frame k + 1), three streams in the batch: every stream bit-identical to the oracle."""
    import vins_fast_b200 as v
    cfg = v.make_config(752, 480, 150, 30, 1, EUROC_CAM0.__dict__, EUROC_CAM1.__dict__)
    cfg.lk_win = win
    S = 3
    batch = v.Batch(cfg, S)
    refs = [oracle.CFeatureTracker(EUROC_CAM0, EUROC_CAM1, 752, 480, 150, 30, 1, 3, win) for _ in range(S)]
    frames = euroc_frames[:8]
    want, keep = [], []
    for k in range(len(frames)):
        # stream s sees the sequence shifted by s frames: three different streams out of one fixture
        imgs = [frames[(k + s) % len(frames)] for s in range(S)]
        want.append([refs[s].track_image(frames[k][0], imgs[s][1], imgs[s][2]) for s in range(S)])
        pl = (C.c_void_p * S)(*[im[1].ctypes.data for im in imgs])
        pr = (C.c_void_p * S)(*[im[2].ctypes.data for im in imgs])
        keep.append((pl, pr))
        batch.enqueue_ptrs(frames[k][0], pl, 752, pr, 752)
        if k >= 1:
            got = batch.fetch(1)
            for s in range(S):
                assert_frames_equal(oracle.records_to_frame(got[s]), want[k - 1][s], f"win {win} lagged frame {k - 1} stream {s}")
    got = batch.fetch(0)
    for s in range(S):
        assert_frames_equal(oracle.records_to_frame(got[s]), want[-1][s], f"win {win} last frame stream {s}")
    batch.close()


def test_mixed_synchronous_and_overlapped_frames(lib, oracle, euroc_frames):
    """The synchronous call and the overlapped enqueue use different stream plans for the same kernels; a sequence may
    switch between them at any frame and must still reproduce the oracle frame after frame (also without CUDA graphs)."""
    import vins_fast_b200 as v
    frames = euroc_frames[:14]
    plan = "ssooosoosssoos"          # s = vf_batch_track, o = vf_batch_enqueue (+ lagged fetch)
    for use_graph in (1, 0):
        cfg = v.make_config(752, 480, 150, 30, 1, EUROC_CAM0.__dict__, EUROC_CAM1.__dict__)
        cfg.use_cuda_graph = use_graph
        batch = v.Batch(cfg, 1)
        ref = oracle.CFeatureTracker(EUROC_CAM0, EUROC_CAM1, 752, 480, 150, 30, 1)
        want, keep, pending = [], [], None
        for k, (t, l, r) in enumerate(frames):
            want.append(ref.track_image(t, l, r))
            pl, pr = (C.c_void_p * 1)(l.ctypes.data), (C.c_void_p * 1)(r.ctypes.data)
            keep.append((pl, pr))
            if plan[k] == "s":
                batch.track_ptrs(t, pl, 752, pr, 752)
                rec = batch._results()[0]
                if pending is not None:      # the overlapped frame before this one is still readable afterwards (lag 1)
                    assert_frames_equal(oracle.records_to_frame(batch.fetch(1)[0]), want[pending], f"graph {use_graph} lagged {pending}")
                    pending = None
                assert_frames_equal(oracle.records_to_frame(rec), want[k], f"graph {use_graph} sync frame {k}")
            else:
                batch.enqueue_ptrs(t, pl, 752, pr, 752)
                if pending is not None:
                    assert_frames_equal(oracle.records_to_frame(batch.fetch(1)[0]), want[pending], f"graph {use_graph} lagged {pending}")
                pending = k
        if pending is not None:
            assert_frames_equal(oracle.records_to_frame(batch.fetch(0)[0]), want[pending], f"graph {use_graph} last")
        batch.close()


def test_long_run_plans_agree_bit_for_bit(lib, euroc_frames):
    """240 frames (the 24 fixture frames played forward and backward): the synchronous plan, the overlapped plan with lag 2
    and a second overlapped run must produce identical record bytes, frame after frame -- the buffers are rings of two /
    three frames, so a missing dependency between streams shows up as a difference within a few hundred frames."""
    import vins_fast_b200 as v
    order = [k if (k // 24) % 2 == 0 else 23 - k % 24 for k in range(240)]
    order = [o % 24 for o in order]
    cfg = v.make_config(752, 480, 150, 30, 1, EUROC_CAM0.__dict__, EUROC_CAM1.__dict__)
    ptrs = [((C.c_void_p * 1)(l.ctypes.data), (C.c_void_p * 1)(r.ctypes.data)) for _, l, r in euroc_frames]

    def run(lag):
        batch = v.Batch(cfg, 1)
        out = []
        for i, k in enumerate(order):
            t = i / 20.0
            if lag < 0:
                batch.track_ptrs(t, ptrs[k][0], 752, ptrs[k][1], 752)
                out.append(batch._results()[0].tobytes())
            else:
                batch.enqueue_ptrs(t, ptrs[k][0], 752, ptrs[k][1], 752)
                if i >= lag:
                    out.append(batch.fetch(lag)[0].tobytes())
        for back in range(lag - 1, -1, -1):
            out.append(batch.fetch(back)[0].tobytes())
        batch.close()
        return out

    want = run(-1)
    assert len(want) == 240 and len(set(want)) > 200
    for name, got in (("lag 2", run(2)), ("lag 2 again", run(2)), ("lag 1", run(1))):
        assert len(got) == 240
        bad = [i for i in range(240) if got[i] != want[i]]
        assert not bad, f"{name}: first differing frame {bad[0]} of {len(bad)}"


def test_4k_2000_features_level5(lib, oracle):
    """BASELINE config 5: 3840x2160 stereo, max_cnt 2000, maxLevel 5 (6 pyramid levels, 16-bit candidate buckets, the
    conflict matrix of the min_dist keep in global memory)."""
    from vins_fast_b200.synth import StereoSequence
    seq = StereoSequence(3, 3840, 2160)
    cam = pinhole_for(3840, 2160)
    _run_pair(oracle, [seq.frame(k) for k in range(3)], 3840, 2160, 2000, 30, 1, cam, cam, level=5)


def test_cpp_frontend_pipeline_feeds_feature_buf(lib, oracle, euroc_frames, tmp_path):
    """estimator::FrontendPipeline (include/vins_fast_b200/frontend_pipeline.hpp): Estimator::FrontendTracker + feature_buf_
    (estimator.cpp:143-158) on the pipelined API; the queued featureFrames equal the synchronous TrackImage results."""
    import struct
    import subprocess
    from conftest import build_cpp_test, write_camera_yaml
    exe = build_cpp_test("test_feature_tracker")
    write_camera_yaml(tmp_path / "cam0_mei.yaml", EUROC_CAM0, 752, 480)
    write_camera_yaml(tmp_path / "cam1_mei.yaml", EUROC_CAM1, 752, 480)
    n = 9
    with open(tmp_path / "frames.bin", "wb") as f:
        for t, l, r in euroc_frames[:n]:
            f.write(struct.pack("<d", t)); f.write(l.tobytes()); f.write(r.tobytes())
    for depth in ("1", "2"):     # frames collected one / two behind the newest
        res = subprocess.run([exe, "pipe", str(tmp_path / "cam0_mei.yaml"), str(tmp_path / "cam1_mei.yaml"),
                              str(tmp_path / "frames.bin"), "752", "480", str(n), depth], capture_output=True, text=True, timeout=300)
        assert res.returncode == 0 and "pipe ok" in res.stdout, res.stdout + res.stderr


def test_device_input_ordering_event_and_caller_stream(lib, oracle, euroc_frames):
    """Device-resident images produced by work still in flight when the frame is enqueued: ordered either by an
    event handed in with vf_batch_set_input_event, or by producing them on the caller-supplied vf_config.cuda_stream."""
    torch = pytest.importorskip("torch")
    import vins_fast_b200 as v
    frames = euroc_frames[:5]
    want = []
    ref = oracle.CFeatureTracker(EUROC_CAM0, EUROC_CAM1, 752, 480, 150, 30, 1)
    for t, l, r in frames:
        want.append(ref.track_image(t, l, r))
    pin = [(torch.from_numpy(l).pin_memory(), torch.from_numpy(r).pin_memory()) for _, l, r in frames]
    big = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")          # a slow producer in front of the images

    # (a) producer on an unrelated stream + event
    prod = torch.cuda.Stream()
    batch = v.Batch(v.make_config(752, 480, 150, 30, 1, EUROC_CAM0.__dict__, EUROC_CAM1.__dict__), 1)
    for k, (t, _, _) in enumerate(frames):
        dl, dr = torch.zeros(480, 752, dtype=torch.uint8, device="cuda"), torch.zeros(480, 752, dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        with torch.cuda.stream(prod):
            big.add_(1)                                                     # ~100 us of work the copies queue behind
            dl.copy_(pin[k][0], non_blocking=True); dr.copy_(pin[k][1], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(prod)
        v.load().vf_batch_set_input_event(batch._h, ev.cuda_event)
        res = batch.track_device(t, dl.data_ptr(), dr.data_ptr(), 752, 752 * 480)
        assert_frames_equal(oracle.records_to_frame(res[0]), want[k], f"event-ordered frame {k}")
    batch.close()

    # (b) the frame work runs on the caller's stream, which also produces the images
    mine = torch.cuda.Stream()
    batch = v.Batch(v.make_config(752, 480, 150, 30, 1, EUROC_CAM0.__dict__, EUROC_CAM1.__dict__, cuda_stream=mine.cuda_stream), 1)
    for k, (t, _, _) in enumerate(frames):
        dl, dr = torch.zeros(480, 752, dtype=torch.uint8, device="cuda"), torch.zeros(480, 752, dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        with torch.cuda.stream(mine):
            big.add_(1)
            dl.copy_(pin[k][0], non_blocking=True); dr.copy_(pin[k][1], non_blocking=True)
        res = batch.track_device(t, dl.data_ptr(), dr.data_ptr(), 752, 752 * 480)
        assert_frames_equal(oracle.records_to_frame(res[0]), want[k], f"caller-stream frame {k}")
    batch.close()


def test_cpp_replay_from_euroc_listing_through_the_pinned_ring(lib, oracle, euroc_frames, tmp_path):
    """EuRoC listing + PGM files -> StereoSync -> FrontendPipeline::Push / AcquireLeft+PushAcquired (pinned
    ingest ring, caller's buffers wiped right after Push) -> feature_buf_; every queued featureFrame equals the oracle's."""
    import struct
    import subprocess
    from conftest import build_cpp_test, write_camera_yaml, write_euroc_dir
    from oracle.cv2_oracle import Cv2FeatureTracker
    exe = build_cpp_test("test_image_source")
    write_camera_yaml(tmp_path / "cam0_mei.yaml", EUROC_CAM0, 752, 480)
    write_camera_yaml(tmp_path / "cam1_mei.yaml", EUROC_CAM1, 752, 480)
    expect = write_euroc_dir(str(tmp_path / "mav0"), euroc_frames[:10], jitter_ns=(0, -1_500_000), drop_right=(4,))
    res = subprocess.run([exe, "replay", str(tmp_path / "mav0"), str(tmp_path / "cam0_mei.yaml"), str(tmp_path / "cam1_mei.yaml"),
                          str(tmp_path / "out.bin")], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0 and "replay ok: 9 pairs, thrown 1 0" in res.stdout, res.stdout + res.stderr
    buf = open(tmp_path / "out.bin", "rb").read()
    off = 0
    ref = oracle.CFeatureTracker(EUROC_CAM0, EUROC_CAM1, 752, 480, 150, 30, 1)
    for k, (t, l, r, ns) in enumerate(expect):
        want = Cv2FeatureTracker.feature_frame(ref.track_image(ns * 1e-9, l, r))     # the reader's stamp: ns * 1e-9
        (gt, n_ids) = struct.unpack_from("<di", buf, off); off += 12
        assert gt == ns * 1e-9 and n_ids == len(want)
        for fid in sorted(want):
            gid, n_obs = struct.unpack_from("<ii", buf, off); off += 8
            assert gid == fid and n_obs == len(want[fid])
            for cam_id, vec in want[fid]:
                (gcam,) = struct.unpack_from("<i", buf, off); off += 4
                v = np.frombuffer(buf, np.float64, 7, off); off += 56
                assert gcam == cam_id and v.tobytes() == vec.tobytes(), f"pair {k} id {fid} cam {cam_id}"
    assert off == len(buf)
