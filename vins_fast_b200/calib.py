"""Calibration constants of the reference's shipped configurations, restated as data, and the synthetic pinhole
used for the 720p / 4K workloads.  Plain dicts in the field layout of vf_camera (include/vins_fast_b200.h).

  EUROC_CAM0 / EUROC_CAM1   vins/config/euroc/cam0_mei.yaml:1-18, cam1_mei.yaml:1-18 (MEI)
  D435_CAM0                 vins/config/fast/cam0_pinhole_d435f.yaml:1-16 (pinhole, zero distortion)
"""
from __future__ import annotations

MEI, PINHOLE = 0, 1


def _cam(model, xi, k1, k2, p1, p2, fx, fy, cx, cy):
    return dict(model=model, xi=xi, k1=k1, k2=k2, p1=p1, p2=p2, fx=fx, fy=fy, cx=cx, cy=cy)


EUROC_CAM0 = _cam(MEI, 3.6313355285286337e+00, 1.1757726639872075e+00, 1.5491281051140213e+01,
                  -8.1237172954550494e-04, 6.7297684030310243e-04,
                  2.1387619122017772e+03, 2.1315886210259278e+03, 3.6119856633263799e+02, 2.4827644773395667e+02)
EUROC_CAM1 = _cam(MEI, 8.1261505894146113e-01, -3.4247886078397966e-01, 1.7108792755973357e-01,
                  -6.6653144364906686e-04, 7.2311255846767091e-04,
                  8.3406249735437791e+02, 8.3140606765916948e+02, 3.7432007355249738e+02, 2.5422391621480082e+02)
D435_CAM0 = _cam(PINHOLE, 0.0, 0.0, 0.0, 0.0, 0.0,
                 388.9324645996094, 388.9324645996094, 323.60009765625, 238.6774444580078)


def pinhole_for(width: int, height: int, distorted: bool = True) -> dict:
    """Synthetic pinhole for the 720p / 4K configurations: f = 0.55 W, mild radtan."""
    f = 0.55 * width
    if distorted:
        return _cam(PINHOLE, 0.0, -0.28, 0.07, 2.0e-4, -1.5e-4, f, f, width / 2.0, height / 2.0)
    return _cam(PINHOLE, 0.0, 0.0, 0.0, 0.0, 0.0, f, f, width / 2.0, height / 2.0)


# BASELINE.json configurations as tracker parameters: "N-level pyramid" = maxLevel N
WORKLOADS = {
    "c1": dict(width=752, height=480, max_cnt=150, min_dist=30, flow_back=1, lk_max_level=3, cams="euroc_mei"),
    "c3": dict(width=1280, height=720, max_cnt=400, min_dist=30, flow_back=1, lk_max_level=4, cams="pinhole"),
    "c5": dict(width=3840, height=2160, max_cnt=2000, min_dist=30, flow_back=1, lk_max_level=5, cams="pinhole"),
}


def workload_cameras(name: str):
    w = WORKLOADS[name]
    if w["cams"] == "euroc_mei":
        return EUROC_CAM0, EUROC_CAM1
    c = pinhole_for(w["width"], w["height"])
    return c, c
