"""batched landmark triangulation and reprojection-error outlier rejection on the GPU against the
REFERENCE'S OWN functions (FeatureManager::TriangulatePts / TriangulateOnePoint / ReProjectionError / OutliersRejection,
vins/estimator/feature_manager.cpp:202-293, 365-429): the committed fixture generated from them
(tests/golden/make_golden_geometry.py) and, where oracle/_ref/libvf_ref_geometry.so is present, the library itself.

Tolerance (float64 work; the null vector of the 4x4 design matrix comes from a one-sided Jacobi SVD here and from Eigen's
two-sided JacobiSVD in the reference): |x - x_ref| <= 1e-10 (1 + |x_ref|) for points and depths; the remove flags must agree
wherever the mean error is not within 1e-9 of the decision threshold."""
import os

import numpy as np
import pytest

from geometry_scene import make_scene, stereo_pairs

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "geometry_reference_eigen_3_3_7.npz")
RTOL = 1e-10


def _close(a, b):
    return np.all(np.abs(a - b) <= RTOL * (1 + np.abs(b)))


@pytest.mark.parametrize("seed,n", [(0, 600), (1, 2500)])
def test_triangulation_and_outlier_rejection_match_the_reference(lib, seed, n):
    from vins_fast_b200 import ops
    g = np.load(GOLD)
    sc = make_scene(seed, n)
    P0, P1, a, b = stereo_pairs(sc)
    assert tuple(g[f"s{seed}_n"]) == (n, len(P0))
    p3 = ops.triangulate_one_point(P0, P1, a, b)
    assert _close(p3, g[f"s{seed}_point3d"]), np.abs(p3 - g[f"s{seed}_point3d"]).max()
    win = (sc["Ps"], sc["Rs"], sc["tic"], sc["ric"])
    tab = (sc["start_frame"], sc["n_obs"], sc["obs_offset"], sc["obs"])
    d = ops.triangulate_pts(*win, *tab, sc["depth"], init_depth=5.0)
    assert _close(d, g[f"s{seed}_depth"])
    untouched = (sc["depth"] > 0) | ((sc["n_obs"] == 1) & (sc["obs"][sc["obs_offset"], 3] == 0))
    assert np.array_equal(d[untouched], sc["depth"][untouched])           # already triangulated / not triangulable: left alone
    ave, rm = ops.outliers_rejection(*win, *tab, g[f"s{seed}_depth"], focal_length=460.0)
    decided = np.abs(ave * 460.0 - 3) > 1e-9
    assert np.array_equal(rm[decided], g[f"s{seed}_remove"][decided])
    assert rm[sc["n_obs"] < 4].sum() == 0 and 0 < rm.sum() < (sc["n_obs"] >= 4).sum()


def test_geometry_against_the_live_reference_library(lib):
    from oracle import ref_geometry as R
    if not R.available():
        pytest.skip("oracle/_ref/libvf_ref_geometry.so not present and /root/reference not available to build it")
    from vins_fast_b200 import ops
    sc = make_scene(7, 1500)
    win = (sc["Ps"], sc["Rs"], sc["tic"], sc["ric"])
    tab = (sc["start_frame"], sc["n_obs"], sc["obs_offset"], sc["obs"])
    d_ref = R.triangulate_pts(*win, *tab, sc["depth"])
    assert _close(ops.triangulate_pts(*win, *tab, sc["depth"]), d_ref)
    ave, rm = ops.outliers_rejection(*win, *tab, d_ref)
    rm_ref = R.outliers_rejection(*win, *tab, d_ref)
    decided = np.abs(ave * 460.0 - 3) > 1e-9
    assert np.array_equal(rm[decided], rm_ref[decided])
    # ReProjectionError itself, through single-observation landmarks is not reachable; check the mean of a 4-frame landmark
    i = int(np.argmax(sc["n_obs"] >= 4))
    Rs, ric = sc["Rs"].reshape(-1, 9), sc["ric"].reshape(2, 9)
    s, m, o = sc["start_frame"][i], sc["n_obs"][i], sc["obs_offset"][i]
    errs = []
    for k in range(m):
        row = sc["obs"][o + k]
        if k > 0:
            errs.append(R.reprojection_error(Rs[s], sc["Ps"][s], ric[0], sc["tic"][0], Rs[s + k], sc["Ps"][s + k], ric[0], sc["tic"][0],
                                             [d_ref[i]], sc["obs"][o, 0:3], row[0:3])[0])
        if row[3]:
            errs.append(R.reprojection_error(Rs[s], sc["Ps"][s], ric[0], sc["tic"][0], Rs[s + k], sc["Ps"][s + k], ric[1], sc["tic"][1],
                                             [d_ref[i]], sc["obs"][o, 0:3], row[4:7])[0])
    assert abs(sum(errs) / len(errs) - ave[i]) <= 1e-12 * (1 + ave[i])


def test_geometry_argument_errors(lib):
    from vins_fast_b200 import VfError, ops
    sc = make_scene(0, 20)
    with pytest.raises(VfError):
        ops.triangulate_pts(sc["Ps"], sc["Rs"], sc["tic"], sc["ric"], sc["start_frame"], sc["n_obs"], sc["obs_offset"] + 10_000, sc["obs"], sc["depth"])
    assert len(ops.triangulate_one_point(np.zeros((0, 12)), np.zeros((0, 12)), np.zeros((0, 2)), np.zeros((0, 2)))) == 0
