"""Pins the CPU oracle (oracle/vf_oracle.c, oracle/camera.py, oracle/stdsort_helper.cpp).

The reference holds no golden vectors for FeatureTracker::TrackImage ("parity unpinned"), and its
arithmetic lives in OpenCV, which is not vendored.  The anchor is therefore OpenCV itself: tests/golden/*.npz were
produced by tests/golden/make_golden.py with the cv2 4.13.0 wheel calling the same routines with the reference's
arguments (feature_tracker.cpp:41-42,50-53,90,117-119,125-127,238,243), and the plain-C restatement must reproduce
them BIT FOR BIT -- pyramid, Scharr, response, corner lists, LK points/status, std::sort permutations and whole
TrackImage sequences.  Where cv2 is importable the same comparison is repeated live on further inputs.
No GPU involved: these tests run in the CPU suite.
"""
import hashlib
import os

import numpy as np
import pytest

from oracle.camera import D435_CAM0, EUROC_CAM0, EUROC_CAM1, pinhole_for

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.fixture(scope="module")
def stages():
    return np.load(os.path.join(GOLD, "stages_cv2_4_13.npz"))


@pytest.fixture(scope="module")
def crops(stages):
    from vins_fast_b200.synth import StereoSequence
    seq = StereoSequence(0, 752, 480)
    _, l0, r0 = seq.frame(0)
    _, l1, _ = seq.frame(1)
    a, b, c = l0[40:200, 100:300].copy(), l1[40:200, 100:300].copy(), r0[40:200, 100:300].copy()
    # the synthetic generator is integer-only: identical bytes on every host, or the fixtures mean nothing
    assert _sha(a) == str(stages["crop_sha_a"]) and _sha(b) == str(stages["crop_sha_b"]) and _sha(c) == str(stages["crop_sha_c"])
    return a, b, c


@pytest.mark.parametrize("tag", ["even", "odd"])
def test_pyramid_and_scharr_golden(oracle, stages, crops, tag):
    img = crops[0] if tag == "even" else crops[0][:159, :199].copy()
    pyr = oracle.Pyramid(img, 3, 21)
    assert pyr.n_levels == int(stages[f"pyr_{tag}_n"])
    for l in range(pyr.n_levels):
        assert np.array_equal(pyr.level(l), stages[f"pyr_{tag}_{l}"]), f"level {l}"
        assert np.array_equal(oracle.scharr(pyr.level(l)), stages[f"scharr_{tag}_{l}"]), f"scharr level {l}"
    assert np.array_equal(oracle.pyrdown(img), stages[f"pyr_{tag}_1"])


def test_min_eigen_golden(oracle, stages, crops):
    eig = oracle.min_eigen_val(crops[0])
    assert np.array_equal(_bits(eig), _bits(stages["eig_noopt"]))       # canonical (non-FMA) order: bit-exact
    # the AVX2/FMA build of OpenCV differs in low bits on some pixels, never by more than a few ulp
    opt = stages["eig_opt"]
    assert np.abs(eig - opt).max() <= 4e-7 * max(1.0, float(np.abs(opt).max()))


def test_good_features_golden(oracle, stages, crops):
    a = crops[0]
    assert np.array_equal(oracle.good_features(a, 60, 0.01, 10)[0], stages["gftt_60_10"])
    assert np.array_equal(oracle.good_features(a, 500, 0.01, 3)[0], stages["gftt_500_3"])
    assert np.array_equal(oracle.good_features(a, 60, 0.01, 10, mask=stages["mask"])[0], stages["gftt_masked"])
    # selection is the same whether the response comes from the canonical or the FMA-contracted OpenCV build
    assert np.array_equal(oracle.good_features(a, 60, 0.01, 10, eig=stages["eig_opt"])[0], stages["gftt_60_10"])


def test_mask_disc_equals_cv_circle(oracle, stages):
    rng = np.random.default_rng(3)
    mask = np.full((160, 200), 255, np.uint8)
    for _ in range(25):
        oracle.mask_disc(mask, int(rng.integers(0, 200)), int(rng.integers(0, 160)), int(rng.integers(0, 40)))
    assert np.array_equal(mask, stages["mask"])


@pytest.mark.skipif(os.environ.get("VF_SKIP_CV2") is not None, reason="cv2 disabled")
def test_mask_disc_equals_cv_circle_over_the_accepted_min_dist_range(oracle):
    """min_dist is accepted up to 1024 (vf_api.cu validate_config); SetFeatureExtractorMask's cv::circle(FILLED) must be the
    Euclidean disc dx^2 + dy^2 <= r^2 over that whole range, clipped centres included (live cv2)."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(11)
    w, h = 1300, 900
    for r in (0, 1, 2, 15, 30, 119, 120, 121, 200, 255, 256, 257, 400, 511, 512, 640, 1000, 1023, 1024):
        for cx, cy in ((w // 2, h // 2), (0, 0), (w - 1, h - 1), (3, h - 2), (int(rng.integers(0, w)), int(rng.integers(0, h)))):
            a = np.full((h, w), 255, np.uint8)
            b = a.copy()
            cv2.circle(a, (cx, cy), r, 0, -1)
            oracle.mask_disc(b, cx, cy, r)
            assert np.array_equal(a, b), f"r={r} centre=({cx},{cy})"


@pytest.mark.parametrize("pair", ["temporal", "stereo"])
def test_lk_golden(oracle, stages, crops, pair):
    a, b, c = crops
    x, y = (a, b) if pair == "temporal" else (a, c)
    pts = stages["lk_pts"]
    px, py = oracle.Pyramid(x), oracle.Pyramid(y)
    nxt, st = oracle.lk(px, py, pts, 3, 21)
    assert np.array_equal(st, stages[f"lk_{pair}_st"])
    ok = st.astype(bool)
    assert ok.sum() > 200
    assert np.array_equal(_bits(nxt[ok]), _bits(stages[f"lk_{pair}_pts"][ok]))
    # points with status 0: OpenCV still reports where the search stopped
    assert np.array_equal(_bits(nxt), _bits(stages[f"lk_{pair}_pts"]))
    rev, rst = oracle.lk(py, px, nxt, 1, 21, init=pts)                  # feature_tracker.cpp:50-53
    assert np.array_equal(rst, stages[f"lk_{pair}_rev_st"])
    assert np.array_equal(_bits(rev), _bits(stages[f"lk_{pair}_rev_pts"]))


@pytest.fixture(scope="module")
def lk_windows():
    return np.load(os.path.join(GOLD, "lk_windows_cv2_4_13.npz"))


def test_lk_golden_other_window_sizes(oracle, lk_windows, crops):
    """`lk_win` other than the reference's literal 21: OpenCV's accumulation order is a function of the window width, so each
    width pins its own chain layout of the C oracle (tests/golden/make_golden_lk_windows.py, cv2 4.13)."""
    a, b, _ = crops
    g = lk_windows
    assert _sha(a) == str(g["crop_sha_a"]) and _sha(b) == str(g["crop_sha_b"])
    pts = g["pts"]
    for win in (int(w) for w in g["windows"]):
        for level in (0, 3):
            nxt, st = oracle.lk(oracle.Pyramid(a, level, win), oracle.Pyramid(b, level, win), pts, level, win)
            assert np.array_equal(st, g[f"w{win}_l{level}_st"]), (win, level)
            assert np.array_equal(_bits(nxt), _bits(g[f"w{win}_l{level}_pts"])), (win, level)
        rev, rst = oracle.lk(oracle.Pyramid(b, 1, win), oracle.Pyramid(a, 1, win), nxt, 1, win, init=pts)
        assert np.array_equal(rst, g[f"w{win}_rev_st"]), win
        assert np.array_equal(_bits(rev), _bits(g[f"w{win}_rev_pts"])), win


@pytest.mark.parametrize("name", ["win15", "win31"])
def test_track_image_golden_other_window_sizes(oracle, lk_windows, name):
    from vins_fast_b200.synth import StereoSequence
    g = lk_windows
    w, h, max_cnt, min_dist, flow_back, level, n_frames, seed, win = (int(v) for v in g[f"{name}_params"])
    cam = pinhole_for(w, h)
    seq = StereoSequence(seed, w, h)
    tr = oracle.CFeatureTracker(cam, cam, w, h, max_cnt, min_dist, flow_back, level, win)
    for k in range(n_frames):
        t, l, r = seq.frame(k)
        f = tr.track_image(t, l, r)
        for key, v in f.items():
            ref = g[f"{name}_f{k}_{key}"]
            assert v.shape == ref.shape, f"frame {k} {key}"
            assert v.tobytes() == ref.astype(v.dtype).tobytes(), f"frame {k} {key}"
        assert len(f["ids"]) > max_cnt // 2


@pytest.mark.parametrize("n", [17, 150, 400])
def test_std_sort_permutation_golden(stages, n):
    from oracle.cv2_oracle import stdsort_perm
    assert np.array_equal(stdsort_perm(stages[f"sort_cnt_{n}"]), stages[f"sort_perm_{n}"])


@pytest.mark.parametrize("name,cams", [
    ("euroc_mei", (EUROC_CAM0, EUROC_CAM1)),
    ("qvga_pinhole", (pinhole_for(320, 240), pinhole_for(320, 240))),
    ("qvga_noflowback", (pinhole_for(320, 240), pinhole_for(320, 240))),
])
def test_track_image_sequences_golden(oracle, name, cams):
    """Whole TrackImage, frame after frame: ids, track counts, pixel / undistorted coordinates, velocities, right
    observations -- every array bit-identical to the cv2 restatement of the reference."""
    from vins_fast_b200.synth import StereoSequence
    g = np.load(os.path.join(GOLD, f"sequence_{name}_cv2_4_13.npz"))
    w, h, max_cnt, min_dist, flow_back, level, n_frames, seed = (int(v) for v in g["params"])
    seq = StereoSequence(seed, w, h)
    tr = oracle.CFeatureTracker(cams[0], cams[1], w, h, max_cnt, min_dist, flow_back, level)
    seen = 0
    for k in range(n_frames):
        t, l, r = seq.frame(k)
        assert np.frombuffer(bytes.fromhex(_sha(l) + _sha(r)), np.uint8).tobytes() == g[f"in_sha_{k}"].tobytes()
        f = tr.track_image(t, l, r)
        for key, v in f.items():
            ref = g[f"f{k}_{key}"]
            assert v.shape == ref.shape, f"frame {k} {key}"
            assert v.tobytes() == ref.astype(v.dtype).tobytes(), f"frame {k} {key}"
        seen += len(f["ids"])
    assert seen >= n_frames * max_cnt * 0.8


# ---- known answers of SURVEY.md section 8c for liftProjective (float64 x/z, y/z) --------------------------------
KATS = [
    (EUROC_CAM0, (100.5, 200.25), (-0.63598498127705982, -0.1174846896296942)),
    (EUROC_CAM0, (376.0, 240.0), (0.032065350809237558, -0.01799004193217792)),
    (EUROC_CAM0, (0.0, 0.0), (-1.1454204151417211, -0.78965396636485363)),
    (EUROC_CAM1, (100.5, 200.25), (-0.67988739918048835, -0.13424034794199594)),
    (D435_CAM0, (100.5, 200.25), (-0.57362169004308439, -0.098802357621566306)),
]


@pytest.mark.parametrize("cam,uv,xy", KATS)
def test_lift_projective_known_answers(oracle, cam, uv, xy):
    x, y = cam.undistort(*uv)
    assert abs(x - xy[0]) < 1e-14 and abs(y - xy[1]) < 1e-14
    cx, cy = oracle.undistort(cam, *uv)                                  # plain-C restatement == Python restatement
    assert (cx, cy) == (x, y)


def test_camera_yaml_reader(tmp_path):
    from oracle.camera import MEI, camera_from_yaml
    p = tmp_path / "cam.yaml"
    p.write_text("%YAML:1.0\n---\nmodel_type: MEI\ncamera_name: cam\nimage_width: 752\nimage_height: 480\n"
                 "mirror_parameters:\n   xi: 2.5\ndistortion_parameters:\n   k1: 0.1\n   k2: -0.2\n   p1: 1e-3\n   p2: 2.0e-3\n"
                 "projection_parameters:\n   gamma1: 900.\n   gamma2: 901.5\n   u0: 370.25\n   v0: 240\n")
    c = camera_from_yaml(str(p))
    assert (c.model, c.xi, c.k1, c.k2, c.p1, c.p2, c.fx, c.fy, c.cx, c.cy) == (MEI, 2.5, 0.1, -0.2, 1e-3, 2e-3, 900.0, 901.5, 370.25, 240.0)


# ---- live comparison against cv2 where it is importable (this container and the GPU box image) -------------------
cv2 = pytest.importorskip("cv2") if os.environ.get("VF_SKIP_CV2") is None else None


@pytest.mark.skipif(cv2 is None, reason="cv2 not importable")
@pytest.mark.parametrize("shape,level", [((240, 320), 3), ((131, 97), 1), ((480, 752), 3)])
def test_oracle_vs_live_cv2_stages(oracle, shape, level):
    from vins_fast_b200.synth import StereoSequence
    h, w = shape
    seq = StereoSequence(21, max(w, 64), max(h, 64))
    _, a, c = seq.frame(3)
    _, b, _ = seq.frame(4)
    a, b, c = (np.ascontiguousarray(x[:h, :w]) for x in (a, b, c))
    _, pyr = cv2.buildOpticalFlowPyramid(a, (21, 21), level, withDerivatives=True)
    op = oracle.Pyramid(a, level, 21)
    assert op.n_levels == len(pyr) // 2
    for l in range(op.n_levels):
        assert np.array_equal(op.level(l), pyr[2 * l])
        assert np.array_equal(oracle.scharr(op.level(l)), pyr[2 * l + 1])
    cv2.setUseOptimized(False)
    try:
        eig = cv2.cornerMinEigenVal(a, 3, ksize=3)
    finally:
        cv2.setUseOptimized(True)
    assert np.array_equal(_bits(oracle.min_eigen_val(a)), _bits(eig))
    for n, md in ((150, 30), (200, 15), (40, 3)):
        ref = cv2.goodFeaturesToTrack(a, n, 0.01, md)
        ref = np.zeros((0, 2), np.float32) if ref is None else ref.reshape(-1, 2)
        assert np.array_equal(oracle.good_features(a, n, 0.01, md)[0], ref)
    pts = oracle.good_features(a, 200, 0.01, 7)[0] + np.float32(0.37)
    for x, y in ((a, b), (a, c)):
        nxt, st, _ = cv2.calcOpticalFlowPyrLK(x, y, pts.reshape(-1, 1, 2), None, winSize=(21, 21), maxLevel=level)
        q, s = oracle.lk(oracle.Pyramid(x, level), oracle.Pyramid(y, level), pts, level, 21)
        assert np.array_equal(s, st.reshape(-1))
        assert np.array_equal(_bits(q), _bits(nxt.reshape(-1, 2)))


@pytest.mark.skipif(cv2 is None, reason="cv2 not importable")
@pytest.mark.parametrize("seed", range(10))
def test_oracle_vs_live_cv2_lk_random_windows(oracle, seed):
    """calcOpticalFlowPyrLK with a random odd window 5..31, random level count, image size, point set (inside, on and outside
    the border) and optionally an initial flow: the C oracle against the live OpenCV routine, bit for bit (positions of lost
    points included).  Pins the width-dependent chain layouts beyond the committed fixture."""
    from vins_fast_b200.synth import StereoSequence
    rng = np.random.default_rng(1000 + seed)
    win = int(rng.choice(np.arange(5, 33, 2)))
    level = int(rng.integers(0, 6))
    h, w = int(rng.integers(70, 260)), int(rng.integers(70, 330))
    seq = StereoSequence(30 + seed, 352, 288)
    _, a, c = seq.frame(int(rng.integers(0, 5)))
    _, b, _ = seq.frame(6)
    y0, x0 = int(rng.integers(0, 288 - h)), int(rng.integers(0, 352 - w))
    a, b, c = (np.ascontiguousarray(x[y0:y0 + h, x0:x0 + w]) for x in (a, b, c))
    pts = np.stack([rng.uniform(-win, w + win, 120), rng.uniform(-win, h + win, 120)], 1).astype(np.float32)
    for x, y in ((a, b), (a, c)):
        nxt, st, _ = cv2.calcOpticalFlowPyrLK(x, y, pts.reshape(-1, 1, 2), None, winSize=(win, win), maxLevel=level)
        q, s = oracle.lk(oracle.Pyramid(x, level, win), oracle.Pyramid(y, level, win), pts, level, win)
        assert np.array_equal(s, st.reshape(-1)), (win, level, h, w)
        assert np.array_equal(_bits(q), _bits(nxt.reshape(-1, 2))), (win, level, h, w)
        lv = min(1, level)
        init = (pts + rng.uniform(-2, 2, pts.shape)).astype(np.float32)
        rev, rst, _ = cv2.calcOpticalFlowPyrLK(y, x, nxt.copy(), init.reshape(-1, 1, 2).copy(), winSize=(win, win), maxLevel=lv,
                                               criteria=(cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 30, 0.01),
                                               flags=cv2.OPTFLOW_USE_INITIAL_FLOW)
        r2, s2 = oracle.lk(oracle.Pyramid(y, lv, win), oracle.Pyramid(x, lv, win), nxt.reshape(-1, 2), lv, win, init=init)
        assert np.array_equal(s2, rst.reshape(-1)), (win, level, h, w)
        assert np.array_equal(_bits(r2), _bits(rev.reshape(-1, 2))), (win, level, h, w)


@pytest.mark.skipif(cv2 is None, reason="cv2 not importable")
def test_oracle_vs_live_cv2_track_image(oracle):
    from oracle.cv2_oracle import Cv2FeatureTracker
    from vins_fast_b200.synth import StereoSequence
    w, h = 376, 240
    cam = pinhole_for(w, h)
    seq = StereoSequence(33, w, h)
    a = Cv2FeatureTracker(cam, cam, 80, 18, 1, 3)
    b = oracle.CFeatureTracker(cam, cam, w, h, 80, 18, 1, 3)
    for k in range(6):
        t, l, r = seq.frame(k)
        fa, fb = a.track_image(t, l, r), b.track_image(t, l, r)
        for key in fa:
            assert np.asarray(fa[key]).tobytes() == np.asarray(fb[key]).astype(np.asarray(fa[key]).dtype).tobytes(), f"frame {k} {key}"


# ---- long sequences (>= 100 frames): one digest per frame, generated by tests/golden/make_golden_long.py with cv2 ------
LONG = [("euroc_mei_120", (EUROC_CAM0, EUROC_CAM1)), ("720p_pinhole_100", (pinhole_for(1280, 720), pinhole_for(1280, 720)))]


@pytest.mark.parametrize("name,cams", LONG)
def test_long_sequences_match_the_cv2_digests(oracle, name, cams):
    """The plain-C oracle against the cv2 restatement of the reference over 120 / 100 consecutive frames (track counts up
    to the sequence length, hundreds of ids issued and dropped): every frame's digest equal."""
    from conftest import frame_digest
    from vins_fast_b200.synth import StereoSequence
    g = np.load(os.path.join(GOLD, f"digest_{name}_cv2_4_13.npz"))
    w, h, max_cnt, min_dist, flow_back, level, n_frames, seed = (int(v) for v in g["params"])
    seq = StereoSequence(seed, w, h)
    tr = oracle.CFeatureTracker(cams[0], cams[1], w, h, max_cnt, min_dist, flow_back, level)
    for k in range(n_frames):
        f = tr.track_image(*seq.frame(k))
        assert (len(f["ids"]), len(f["ids_right"])) == tuple(int(v) for v in g["counts"][k, :2]), f"first divergence at frame {k} (counts)"
        assert frame_digest(f) == g["digest"][k].tobytes(), f"first divergence at frame {k}"
    assert int(g["counts"][:, 2].max()) >= n_frames // 2 and int(g["counts"][-1, 3]) > 2 * max_cnt


def test_vectorised_glue_of_the_cv2_oracle_is_bit_identical():
    """oracle/cv2_oracle.py vectorises the reference's per-point glue with numpy; the per-point forms must give the same bits."""
    rng = np.random.default_rng(5)
    for cam in (EUROC_CAM0, EUROC_CAM1, D435_CAM0, pinhole_for(1280, 720)):
        uv = (rng.uniform(-20, 780, (300, 2))).astype(np.float32)
        many = cam.undistort_many(uv)
        one = np.array([[np.float32(v) for v in cam.undistort(float(p[0]), float(p[1]))] for p in uv], np.float32)
        assert many.tobytes() == one.tobytes()
    from oracle.cv2_oracle import _in_border, _pt_distance
    p = (rng.uniform(-3, 760, (500, 2))).astype(np.float32)
    p[:8] = [[0.5, 0.5], [1.5, 2.5], [750.5, 478.5], [749.5, 477.5], [1.0, 1.0], [750.0, 478.0], [751.0, 479.0], [0.4999, 1.5]]
    q = (p + rng.uniform(-0.6, 0.6, p.shape)).astype(np.float32)
    for i in range(len(p)):
        x, y = int(np.rint(p[i, 0])), int(np.rint(p[i, 1]))
        assert bool(_in_border(p, 480, 752)[i]) == (1 <= x < 751 and 1 <= y < 479)
        dx, dy = float(p[i, 0] - q[i, 0]), float(p[i, 1] - q[i, 1])
        assert float(_pt_distance(p, q)[i]) == float(np.sqrt(dx * dx + dy * dy))


def test_geometry_fixture_is_what_the_reference_code_computes():
    """tests/golden/geometry_reference_eigen_3_3_7.npz against the reference's own FeatureManager functions (oracle/_ref, built
    from /root/reference by oracle/build_ref.py); skipped where neither the reference nor the prebuilt library exists."""
    from oracle import ref_geometry as R
    if not R.available():
        pytest.skip("reference not available")
    from geometry_scene import make_scene, stereo_pairs
    g = np.load(os.path.join(GOLD, "geometry_reference_eigen_3_3_7.npz"))
    sc = make_scene(0, 600)
    P0, P1, a, b = stereo_pairs(sc)
    assert np.array_equal(R.triangulate_one_point(P0, P1, a, b), g["s0_point3d"])
    d = R.triangulate_pts(sc["Ps"], sc["Rs"], sc["tic"], sc["ric"], sc["start_frame"], sc["n_obs"], sc["obs_offset"], sc["obs"], sc["depth"])
    assert np.array_equal(d, g["s0_depth"])
    rm = R.outliers_rejection(sc["Ps"], sc["Rs"], sc["tic"], sc["ric"], sc["start_frame"], sc["n_obs"], sc["obs_offset"], sc["obs"], d)
    assert np.array_equal(rm, g["s0_remove"])
    # sanity of the scene: triangulated depths of clean stereo landmarks are the true depths to ~1 %
    assert np.median(d[d > 0]) > 1.0
