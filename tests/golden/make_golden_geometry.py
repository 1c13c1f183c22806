#!/usr/bin/env python
"""Known answers for the triangulation / outlier-rejection operators, produced by the REFERENCE'S OWN code
(FeatureManager::TriangulatePts / TriangulateOnePoint / ReProjectionError / OutliersRejection,
vins/estimator/feature_manager.cpp:202-293, 365-429, compiled from /root/reference against its vendored Eigen 3.3.7 by
oracle/build_ref.py) on the deterministic scene of tests/geometry_scene.py.  The fixture lets the GPU box -- where
/root/reference does not exist -- check against the reference even without the prebuilt oracle/_ref library.
    python tests/golden/make_golden_geometry.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from geometry_scene import make_scene, stereo_pairs  # noqa: E402
from oracle import ref_geometry as R  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

if __name__ == "__main__":
    out = {}
    for seed, n in ((0, 600), (1, 2500)):
        sc = make_scene(seed, n)
        P0, P1, a, b = stereo_pairs(sc)
        out[f"s{seed}_point3d"] = R.triangulate_one_point(P0, P1, a, b)
        d = R.triangulate_pts(sc["Ps"], sc["Rs"], sc["tic"], sc["ric"], sc["start_frame"], sc["n_obs"], sc["obs_offset"], sc["obs"], sc["depth"])
        out[f"s{seed}_depth"] = d
        out[f"s{seed}_remove"] = R.outliers_rejection(sc["Ps"], sc["Rs"], sc["tic"], sc["ric"], sc["start_frame"], sc["n_obs"],
                                                      sc["obs_offset"], sc["obs"], d, 460.0)
        out[f"s{seed}_n"] = np.array([n, len(P0)])
        print(seed, n, "stereo pairs", len(P0), "removed", int(out[f"s{seed}_remove"].sum()))
    np.savez_compressed(os.path.join(OUT, "geometry_reference_eigen_3_3_7.npz"), **out)
