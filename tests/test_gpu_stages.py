"""Stage-level parity: every sm_100a kernel against the CPU oracle (oracle/vf_oracle.c, itself pinned bit for bit
against cv2 4.13 by tests/test_oracle_golden.py), through the C ABI's vf_op_* entry points.
Integer stages, the float response map and the LK results must be BIT-EXACT."""
import numpy as np
import pytest

from oracle.camera import D435_CAM0, EUROC_CAM0, EUROC_CAM1, pinhole_for

pytestmark = pytest.mark.gpu


def _crop(img, h, w, y0=3, x0=5):
    return np.ascontiguousarray(img[y0:y0 + h, x0:x0 + w])


@pytest.fixture(params=[1, 2], ids=["fused", "tma"])
def pyramid_kernel(request, lib):
    """Both pyramid kernels (fused multi-level / streaming TMA; the library picks by size) must give the same bytes."""
    from vins_fast_b200 import ops
    ops.set_pyramid_kernel(request.param)
    yield request.param
    ops.set_pyramid_kernel(0)


@pytest.mark.parametrize("shape", [(480, 752), (479, 751), (240, 376), (97, 131), (64, 64), (720, 1280), (193, 385), (200, 191), (1083, 1925)])
def test_pyramid_bit_exact(lib, oracle, euroc_seq, shape, pyramid_kernel):
    from vins_fast_b200 import ops
    from vins_fast_b200.synth import StereoSequence
    h, w = shape
    seq = euroc_seq if (h <= 480 and w <= 752) else StereoSequence(1, w, h)
    img = _crop(seq.frame(2)[1], h, w) if seq is euroc_seq else seq.frame(2)[1]
    got = ops.build_pyramid(img, 4)
    ref = oracle.Pyramid(img, 4, 21)
    assert len(got) == ref.n_levels
    for l, g in enumerate(got):
        assert np.array_equal(g, ref.level(l)), f"pyramid level {l} differs"


def test_pyramid_early_stop(lib, oracle, euroc_seq):
    from vins_fast_b200 import ops
    img = _crop(euroc_seq.frame(0)[1], 100, 120)       # level 2 would be 30x25 -> next <= 21: stop at level 2
    got = ops.build_pyramid(img, 5)
    assert len(got) == oracle.Pyramid(img, 5, 21).n_levels


@pytest.mark.parametrize("shape", [(480, 752), (61, 95), (64, 64)])
def test_scharr_bit_exact(lib, oracle, euroc_seq, shape):
    from vins_fast_b200 import ops
    img = _crop(euroc_seq.frame(1)[2], *shape)
    assert np.array_equal(ops.scharr(img), oracle.scharr(img))


def _lk_points(img, n, seed):
    rng = np.random.default_rng(seed)
    h, w = img.shape
    pts = np.stack([rng.uniform(-3, w + 3, n), rng.uniform(-3, h + 3, n)], 1).astype(np.float32)
    edge = np.array([[0.3, 0.2], [w - 1, h - 1], [2.5, h / 2], [w - 2.3, 100.2], [w / 2, 1.2], [w / 2, h - 1.4],
                     [-5, -5], [w + 8, 300], [10.0, 10.0], [w - 11.0, h - 11.0]], np.float32)
    return np.concatenate([pts, edge])


@pytest.fixture(params=[2, 3, 1], ids=["warp", "coop", "cta"])
def lk_kernel(request, lib):
    """All Lucas-Kanade kernels (warp-level passes with one warp / one cooperating CTA per feature -- the library picks by launch
    size -- and the block-barrier kernels of vf_lk.cu) must give the same bits."""
    from vins_fast_b200 import ops
    ops.set_lk_kernel(request.param)
    yield request.param
    ops.set_lk_kernel(0)


@pytest.mark.parametrize("pair", ["temporal", "stereo"])
def test_lk_forward_bit_exact(lib, oracle, euroc_frames, pair, lk_kernel):
    from vins_fast_b200 import ops
    _, l0, r0 = euroc_frames[0]
    _, l1, _ = euroc_frames[1]
    a, b = (l0, l1) if pair == "temporal" else (l0, r0)
    pts = _lk_points(a, 600, 7)
    q, st, it = ops.lk(a, b, pts, 3, return_iters=True)
    rq, rst, rit = oracle.lk(oracle.Pyramid(a), oracle.Pyramid(b), pts, 3, 21, return_iters=True)
    assert np.array_equal(st, rst)
    assert np.array_equal(it, rit)
    assert np.array_equal(q.view(np.uint32), rq.view(np.uint32)), np.abs(q - rq).max()
    assert st.sum() > 400


def test_lk_reverse_initial_flow_bit_exact(lib, oracle, euroc_frames, lk_kernel):
    from vins_fast_b200 import ops
    _, l0, _ = euroc_frames[0]
    _, l1, _ = euroc_frames[1]
    pts = _lk_points(l0, 400, 11)
    fwd, st = ops.lk(l0, l1, pts, 3)
    rev, rst = ops.lk(l1, l0, fwd, 1, init=pts)                       # feature_tracker.cpp:50-53
    orev, orst = oracle.lk(oracle.Pyramid(l1), oracle.Pyramid(l0), fwd, 1, 21, init=pts)
    assert np.array_equal(rst, orst)
    assert np.array_equal(rev.view(np.uint32), orev.view(np.uint32))


@pytest.mark.parametrize("level", [0, 1, 4])
def test_lk_levels_and_small_images(lib, oracle, euroc_frames, level, lk_kernel):
    from vins_fast_b200 import ops
    a = _crop(euroc_frames[3][1], 200, 260)
    b = _crop(euroc_frames[4][1], 200, 260)
    pts = _lk_points(a, 200, 3)
    q, st = ops.lk(a, b, pts, level)
    rq, rst = oracle.lk(oracle.Pyramid(a, level), oracle.Pyramid(b, level), pts, level, 21)
    assert np.array_equal(st, rst)
    assert np.array_equal(q.view(np.uint32), rq.view(np.uint32))


def test_lk_six_levels_many_points(lib, oracle, lk_kernel):
    """Six pyramid levels and more points than fit on the GPU at once: the launch takes the variant of the prologue
    that forms the terms of A on the fly instead of staging them in shared memory (vf_lk.cu: lk_plan_staged)."""
    from vins_fast_b200 import ops
    from vins_fast_b200.synth import StereoSequence
    seq = StereoSequence(4, 1024, 768)
    a, b = seq.frame(0)[1], seq.frame(1)[1]
    pts = _lk_points(a, 500, 13)
    q, st, it = ops.lk(a, b, pts, 5, return_iters=True)
    rq, rst, rit = oracle.lk(oracle.Pyramid(a, 5), oracle.Pyramid(b, 5), pts, 5, 21, return_iters=True)
    assert np.array_equal(st, rst)
    assert np.array_equal(it, rit)
    assert np.array_equal(q.view(np.uint32), rq.view(np.uint32)), np.abs(q - rq).max()
    assert st.sum() > 300


def test_lk_textureless_and_empty(lib, oracle, lk_kernel):
    from vins_fast_b200 import ops
    flat = np.full((120, 160), 77, np.uint8)
    pts = np.array([[50, 50], [80.5, 60.25]], np.float32)
    q, st = ops.lk(flat, flat, pts, 3)
    rq, rst = oracle.lk(oracle.Pyramid(flat), oracle.Pyramid(flat), pts, 3, 21)
    assert np.array_equal(st, rst) and st.sum() == 0                    # minEig < 1e-4 -> status 0
    assert np.array_equal(q.view(np.uint32), rq.view(np.uint32))
    q, st = ops.lk(flat, flat, np.zeros((0, 2), np.float32), 3)
    assert q.shape == (0, 2) and st.shape == (0,)


def test_lk_high_contrast_takes_the_ordered_chains(lib, oracle, lk_kernel):
    """A black / white checker moved by half a pixel: |J - I| |grad| sums far beyond 2^24, so the iterations leave the
    exact-integer shortcut and run OpenCV's ordered float chains; positions, status and iteration counts still bit-equal."""
    from vins_fast_b200 import ops
    yy, xx = np.mgrid[0:240, 0:320]
    a = (((xx // 5 + yy // 5) & 1) * 255).astype(np.uint8)
    b = np.roll(a, (1, 2), (0, 1))
    b[::2] = (b[::2].astype(np.int32) * 3 // 4).astype(np.uint8)
    pts = _lk_points(a, 300, 5)
    q, st, it = ops.lk(a, b, pts, 3, return_iters=True)
    rq, rst, rit = oracle.lk(oracle.Pyramid(a), oracle.Pyramid(b), pts, 3, 21, return_iters=True)
    assert np.array_equal(st, rst) and np.array_equal(it, rit)
    assert np.array_equal(q.view(np.uint32), rq.view(np.uint32)), np.abs(q - rq).max()


@pytest.mark.parametrize("win", [5, 7, 9, 11, 13, 15, 17, 19, 23, 25, 27, 29, 31])
def test_lk_other_window_sizes_bit_exact(lib, oracle, euroc_frames, win):
    """calcOpticalFlowPyrLK with winSize != (21, 21) (csrc/vf_lk_generic.cu): forward at 0 / 3 / 5 requested levels (the
    pyramid's early stop follows the window), reverse with an initial flow, points outside the image included --
    positions, status and iteration counts bit-equal to the oracle for that width."""
    from vins_fast_b200 import ops
    _, l0, r0 = euroc_frames[0]
    _, l1, _ = euroc_frames[1]
    pts = np.concatenate([_lk_points(l0, 300, win), np.array([[-40, 50], [760, 470], [0.2, 0.3], [751, 479], [-3, 200]], np.float32)])
    for a, b, level in ((l0, l1, 3), (l0, r0, 0), (_crop(l0, 200, 260), _crop(l1, 200, 260), 5)):
        q, st, it = ops.lk(a, b, pts, level, return_iters=True, win=win)
        rq, rst, rit = oracle.lk(oracle.Pyramid(a, level, win), oracle.Pyramid(b, level, win), pts, level, win, return_iters=True)
        assert np.array_equal(st, rst), level
        assert np.array_equal(it, rit), level
        assert np.array_equal(q.view(np.uint32), rq.view(np.uint32)), (level, np.abs(q - rq).max())
    fwd, _ = ops.lk(l0, l1, pts, 3, win=win)
    rev, rst = ops.lk(l1, l0, fwd, 1, init=pts, win=win)
    orev, orst = oracle.lk(oracle.Pyramid(l1, 1, win), oracle.Pyramid(l0, 1, win), fwd, 1, win, init=pts)
    assert np.array_equal(rst, orst) and np.array_equal(rev.view(np.uint32), orev.view(np.uint32))


@pytest.mark.parametrize("seed", range(10))
def test_lk_random_windows_levels_shapes(lib, oracle, seed):
    """Random odd window 5..31 (21 included: the tuned kernels), level count, image size and point set (inside, on and outside
    the border), forward and reverse with an initial flow -- the same generator the oracle is checked with against live cv2
    (tests/test_oracle_golden.py::test_oracle_vs_live_cv2_lk_random_windows)."""
    from vins_fast_b200 import ops
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
        q, st, it = ops.lk(x, y, pts, level, return_iters=True, win=win)
        rq, rst, rit = oracle.lk(oracle.Pyramid(x, level, win), oracle.Pyramid(y, level, win), pts, level, win, return_iters=True)
        assert np.array_equal(st, rst) and np.array_equal(it, rit), (win, level, h, w)
        assert np.array_equal(q.view(np.uint32), rq.view(np.uint32)), (win, level, h, w)
        lv = min(1, level)
        init = (pts + rng.uniform(-2, 2, pts.shape)).astype(np.float32)
        r2, s2 = ops.lk(y, x, q, lv, init=init, win=win)
        o2, os2 = oracle.lk(oracle.Pyramid(y, lv, win), oracle.Pyramid(x, lv, win), q, lv, win, init=init)
        assert np.array_equal(s2, os2) and np.array_equal(r2.view(np.uint32), o2.view(np.uint32)), (win, level, h, w)


@pytest.mark.parametrize("win", [7, 15, 31])
def test_lk_other_window_sizes_ordered_chains_and_textureless(lib, oracle, win):
    """The high-contrast checker of the 21x21 test (sums beyond 2^24: ordered float chains in the layout of this width) and a
    flat image (minEig below threshold)."""
    from vins_fast_b200 import ops
    yy, xx = np.mgrid[0:240, 0:320]
    a = (((xx // 5 + yy // 5) & 1) * 255).astype(np.uint8)
    b = np.roll(a, (1, 2), (0, 1))
    b[::2] = (b[::2].astype(np.int32) * 3 // 4).astype(np.uint8)
    pts = _lk_points(a, 300, 5)
    q, st, it = ops.lk(a, b, pts, 3, return_iters=True, win=win)
    rq, rst, rit = oracle.lk(oracle.Pyramid(a, 3, win), oracle.Pyramid(b, 3, win), pts, 3, win, return_iters=True)
    assert np.array_equal(st, rst) and np.array_equal(it, rit)
    assert np.array_equal(q.view(np.uint32), rq.view(np.uint32)), np.abs(q - rq).max()
    flat = np.full((120, 160), 77, np.uint8)
    p2 = np.array([[50, 50], [80.5, 60.25]], np.float32)
    q, st = ops.lk(flat, flat, p2, 3, win=win)
    rq, rst = oracle.lk(oracle.Pyramid(flat, 3, win), oracle.Pyramid(flat, 3, win), p2, 3, win)
    assert np.array_equal(st, rst) and st.sum() == 0 and np.array_equal(q.view(np.uint32), rq.view(np.uint32))


@pytest.fixture(params=[1, 2], ids=["tiles", "strips"])
def response_kernel(request, lib):
    """Both corner-response kernels (the library picks by frame size) must give the same bits."""
    from vins_fast_b200 import ops
    ops.set_response_kernel(request.param)
    yield request.param
    ops.set_response_kernel(0)


@pytest.mark.parametrize("shape", [(480, 752), (99, 131), (64, 67), (65, 120), (64, 121), (70, 244), (130, 64), (97, 362)])
def test_min_eigen_bit_exact(lib, oracle, euroc_seq, shape, response_kernel):
    from vins_fast_b200 import ops
    img = _crop(euroc_seq.frame(5)[1], *shape)
    got, ref = ops.min_eigen(img), oracle.min_eigen_val(img)
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32)), (got != ref).mean()


def test_min_eigen_flat_and_saturated(lib, oracle, euroc_seq, response_kernel):
    """Constant patches make the radicand of the eigenvalue exactly 0 (the square root's special-case path)."""
    from vins_fast_b200 import ops
    img = _crop(euroc_seq.frame(3)[1], 200, 360).copy()
    img[40:90, 100:260] = 255
    img[120:124, :] = 0
    img[:, 300:] = 17
    got, ref = ops.min_eigen(img), oracle.min_eigen_val(img)
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32)), (got != ref).mean()


@pytest.mark.parametrize("n,md", [(150, 30), (200, 15), (2000, 30), (50, 7), (400, 1), (3000, 3)])
def test_good_features_selection_order(lib, oracle, euroc_frames, n, md, response_kernel):
    from vins_fast_b200 import ops
    img = euroc_frames[2][1]
    got = ops.good_features(img, n, 0.01, md)
    ref, _ = oracle.good_features(img, n, 0.01, md)
    assert np.array_equal(got, ref)


def test_good_features_masked(lib, oracle, euroc_frames, response_kernel):
    from vins_fast_b200 import ops
    img = euroc_frames[6][1]
    rng = np.random.default_rng(5)
    mask = np.full(img.shape, 255, np.uint8)
    for _ in range(120):
        oracle.mask_disc(mask, int(rng.integers(0, 752)), int(rng.integers(0, 480)), 30)
    got = ops.good_features(img, 150, 0.01, 30, mask=mask)
    ref, _ = oracle.good_features(img, 150, 0.01, 30, mask=mask)
    assert np.array_equal(got, ref) and len(ref) > 0
    empty = np.zeros(img.shape, np.uint8)
    assert len(ops.good_features(img, 10, 0.01, 30, mask=empty)) == 0


def test_good_features_flat_and_plateau(lib, oracle, response_kernel):
    from vins_fast_b200 import ops
    flat = np.full((96, 128), 10, np.uint8)
    assert len(ops.good_features(flat, 50, 0.01, 5)) == 0
    # checkerboard: many exactly equal responses -> ties resolved by raster index (descending), one oversized bucket
    yy, xx = np.mgrid[0:256, 0:320]
    board = (((yy // 8) + (xx // 8)) % 2 * 200 + 20).astype(np.uint8)
    got = ops.good_features(board, 500, 0.01, 4)
    ref, nc = oracle.good_features(board, 500, 0.01, 4)
    assert nc > 1100
    assert np.array_equal(got, ref)


def _ref_set_mask(oracle, pts, cnt, w, h, md):
    from oracle.cv2_oracle import stdsort_perm
    perm = stdsort_perm(cnt)
    mask = np.full((h, w), 255, np.uint8)
    keep = []
    for j in perm:
        cx, cy = int(np.rint(pts[j, 0])), int(np.rint(pts[j, 1]))
        if mask[cy, cx] == 255:
            keep.append(j)
            oracle.mask_disc(mask, cx, cy, md)
    return perm, np.array(keep, np.int32), mask


@pytest.mark.parametrize("n,md,seed", [(150, 30, 0), (200, 15, 1), (17, 30, 2), (16, 30, 3), (1, 30, 4), (700, 10, 5),
                                       (2000, 30, 6), (300, 0, 7)])
def test_set_mask_matches_std_sort_and_greedy(lib, oracle, n, md, seed):
    from vins_fast_b200 import ops
    rng = np.random.default_rng(seed)
    w, h = 752, 480
    pts = np.stack([rng.uniform(1, w - 2, n), rng.uniform(1, h - 2, n)], 1).astype(np.float32)
    cnt = np.sort(rng.integers(1, 12, n)).astype(np.int32)[::-1].copy()      # tracker-like: long runs of equal counts
    cnt[rng.integers(0, n, max(n // 10, 1))] += 3
    perm, keep, mask = ops.set_mask(pts, cnt, w, h, md)
    rperm, rkeep, rmask = _ref_set_mask(oracle, pts, cnt, w, h, md)
    assert np.array_equal(perm, rperm)
    assert np.array_equal(keep, rkeep)
    assert np.array_equal(mask, rmask)


def test_set_mask_heap_sort_fallback(lib, oracle):
    """Median-of-3 killer keys push libstdc++'s introsort to its depth limit (heap-sort path)."""
    from vins_fast_b200 import ops
    n = 2048
    half = n // 2
    k = np.zeros(n, np.int64)
    for i in range(half):
        k[i] = i + 1 if i % 2 == 0 else half + i
        k[half + i] = 2 * (i + 1)
    cnt = (3 * n - k).astype(np.int32)
    rng = np.random.default_rng(0)
    pts = np.stack([rng.uniform(1, 750, n), rng.uniform(1, 478, n)], 1).astype(np.float32)
    perm, keep, _ = ops.set_mask(pts, cnt, 752, 480, 5)
    rperm, rkeep, _ = _ref_set_mask(oracle, pts, cnt, 752, 480, 5)
    assert np.array_equal(perm, rperm) and np.array_equal(keep, rkeep)


@pytest.mark.parametrize("cam", [EUROC_CAM0, EUROC_CAM1, D435_CAM0, pinhole_for(1280, 720)])
def test_undistort_bit_exact(lib, oracle, cam):
    from vins_fast_b200 import ops
    rng = np.random.default_rng(9)
    uv = np.stack([rng.uniform(0, 752, 500), rng.uniform(0, 480, 500)], 1).astype(np.float32)
    uv[:3] = [[100.5, 200.25], [376, 240], [0, 0]]
    got = ops.undistort(cam.__dict__, uv)
    ref = np.array([[np.float32(v) for v in oracle.undistort(cam, float(u[0]), float(u[1]))] for u in uv], np.float32)
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32)), np.abs(got - ref).max()
