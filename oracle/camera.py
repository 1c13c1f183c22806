"""TEST INFRASTRUCTURE (oracle) -- float64 restatement of camodocal liftProjective.

Follows, expression by expression (same association order, no FMA):
  * CataCamera::liftProjective   camera_models/src/camera_models/CataCamera.cc:556-626
  * CataCamera::distortion       camera_models/src/camera_models/CataCamera.cc:766-782
  * inverse-K parameters         camera_models/src/camera_models/CataCamera.cc:320-323
  * PinholeCamera::liftProjective camera_models/src/camera_models/PinholeCamera.cc:450-510
  * PinholeCamera::distortion    camera_models/src/camera_models/PinholeCamera.cc:646-662
  * YAML keys                    CataCamera.cc:161-202, PinholeCamera.cc:145-183,
                                 CameraFactory.cc:96-136 (model_type, case-insensitive)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference leg may import this.
Pinned by the known answers of SURVEY.md §8c (tests/test_oracle_golden.py::test_lift_projective_known_answers).
"""
from __future__ import annotations

import math
import re
from dataclasses import dataclass

MEI, PINHOLE = 0, 1


@dataclass
class Camera:
    model: int
    xi: float
    k1: float
    k2: float
    p1: float
    p2: float
    fx: float  # gamma1 for MEI
    fy: float  # gamma2 for MEI
    cx: float  # u0 for MEI
    cy: float  # v0 for MEI

    def inv_k(self):
        return 1.0 / self.fx, -self.cx / self.fx, 1.0 / self.fy, -self.cy / self.fy

    @property
    def no_distortion(self) -> bool:
        return self.k1 == 0.0 and self.k2 == 0.0 and self.p1 == 0.0 and self.p2 == 0.0

    def distortion(self, x: float, y: float):
        k1, k2, p1, p2 = self.k1, self.k2, self.p1, self.p2
        mx2 = x * x
        my2 = y * y
        mxy = x * y
        rho2 = mx2 + my2
        rad = k1 * rho2 + k2 * rho2 * rho2
        dx = x * rad + 2.0 * p1 * mxy + p2 * (rho2 + 2.0 * mx2)
        dy = y * rad + 2.0 * p2 * mxy + p1 * (rho2 + 2.0 * my2)
        return dx, dy

    def lift_projective(self, u: float, v: float):
        i11, i13, i22, i23 = self.inv_k()
        mx_d = i11 * u + i13
        my_d = i22 * v + i23
        if self.no_distortion:
            mx_u, my_u = mx_d, my_d
        else:
            dx, dy = self.distortion(mx_d, my_d)
            mx_u, my_u = mx_d - dx, my_d - dy
            for _ in range(1, 8):
                dx, dy = self.distortion(mx_u, my_u)
                mx_u, my_u = mx_d - dx, my_d - dy
        if self.model == PINHOLE:
            return mx_u, my_u, 1.0
        xi = self.xi
        if xi == 1.0:
            return mx_u, my_u, (1.0 - mx_u * mx_u - my_u * my_u) / 2.0
        rho2 = mx_u * mx_u + my_u * my_u
        z = 1.0 - xi * (rho2 + 1.0) / (xi + math.sqrt(1.0 + (1.0 - xi * xi) * rho2))
        return mx_u, my_u, z

    def undistort(self, u: float, v: float):
        """RemoveDistortion (feature_tracker.cpp:310-321): (x/z, y/z) in float64 (caller casts to f32)."""
        x, y, z = self.lift_projective(float(u), float(v))
        return x / z, y / z

    def undistort_many(self, uv):
        """RemoveDistortion for an [N,2] float32 array -> [N,2] float32.  The same float64 expressions as
        lift_projective / distortion, evaluated element-wise by numpy in the same association order (no FMA), so every
        element is bit-identical to the scalar path (tests/test_oracle_golden.py checks it)."""
        import numpy as np
        uv = np.asarray(uv, np.float32).reshape(-1, 2)
        u = uv[:, 0].astype(np.float64)
        v = uv[:, 1].astype(np.float64)
        i11, i13, i22, i23 = self.inv_k()
        mx_d = i11 * u + i13
        my_d = i22 * v + i23
        if self.no_distortion:
            mx_u, my_u = mx_d, my_d
        else:
            dx, dy = self.distortion(mx_d, my_d)
            mx_u, my_u = mx_d - dx, my_d - dy
            for _ in range(1, 8):
                dx, dy = self.distortion(mx_u, my_u)
                mx_u, my_u = mx_d - dx, my_d - dy
        if self.model == PINHOLE:
            z = np.ones_like(mx_u)
        elif self.xi == 1.0:
            z = (1.0 - mx_u * mx_u - my_u * my_u) / 2.0
        else:
            xi = self.xi
            rho2 = mx_u * mx_u + my_u * my_u
            z = 1.0 - xi * (rho2 + 1.0) / (xi + np.sqrt(1.0 + (1.0 - xi * xi) * rho2))
        return np.stack([(mx_u / z).astype(np.float32), (my_u / z).astype(np.float32)], 1)


def parse_opencv_yaml_scalars(text: str) -> dict:
    """Tiny reader for the two-level scalar maps of the camodocal calibration files."""
    out: dict = {}
    section = None
    for raw in text.splitlines():
        line = raw.split("#", 1)[0].rstrip()
        if not line or line.startswith("%") or line.strip() == "---":
            continue
        m = re.match(r"^(\s*)([A-Za-z_0-9]+)\s*:\s*(.*)$", line)
        if not m:
            continue
        indent, key, val = len(m.group(1)), m.group(2), m.group(3).strip().strip('"')
        if indent == 0:
            if val == "":
                section = key
                out[key] = {}
            else:
                section = None
                out[key] = val
        elif section is not None:
            out[section][key] = val
    return out


def camera_from_yaml(path: str) -> Camera:
    with open(path) as f:
        d = parse_opencv_yaml_scalars(f.read())
    model = str(d.get("model_type", "MEI")).lower()
    dist = {k: float(v) for k, v in d.get("distortion_parameters", {}).items()}
    proj = {k: float(v) for k, v in d.get("projection_parameters", {}).items()}
    k = [dist.get(n, 0.0) for n in ("k1", "k2", "p1", "p2")]
    if model == "pinhole":
        return Camera(PINHOLE, 0.0, *k, proj["fx"], proj["fy"], proj["cx"], proj["cy"])
    if model == "mei":
        xi = float(d.get("mirror_parameters", {}).get("xi", 0.0))
        return Camera(MEI, xi, *k, proj["gamma1"], proj["gamma2"], proj["u0"], proj["v0"])
    raise ValueError(f"unsupported camera model {model!r} (hot path covers MEI and PINHOLE)")


# EuRoC calibration of the reference (vins/config/euroc/cam0_mei.yaml:1-18, cam1_mei.yaml:1-18) and the
# D435 pinhole (vins/config/fast/cam0_pinhole_d435f.yaml:1-16); numeric parameters, restated as data.
EUROC_CAM0 = Camera(MEI, 3.6313355285286337e+00, 1.1757726639872075e+00, 1.5491281051140213e+01,
                    -8.1237172954550494e-04, 6.7297684030310243e-04,
                    2.1387619122017772e+03, 2.1315886210259278e+03, 3.6119856633263799e+02, 2.4827644773395667e+02)
EUROC_CAM1 = Camera(MEI, 8.1261505894146113e-01, -3.4247886078397966e-01, 1.7108792755973357e-01,
                    -6.6653144364906686e-04, 7.2311255846767091e-04,
                    8.3406249735437791e+02, 8.3140606765916948e+02, 3.7432007355249738e+02, 2.5422391621480082e+02)
D435_CAM0 = Camera(PINHOLE, 0.0, 0.0, 0.0, 0.0, 0.0,
                   388.9324645996094, 388.9324645996094, 323.60009765625, 238.6774444580078)


def pinhole_for(width: int, height: int, distorted: bool = True) -> Camera:
    """Synthetic pinhole for the 720p / 4K configs: f = 0.55*W, mild radtan."""
    f = 0.55 * width
    if distorted:
        return Camera(PINHOLE, 0.0, -0.28, 0.07, 2.0e-4, -1.5e-4, f, f, width / 2.0, height / 2.0)
    return Camera(PINHOLE, 0.0, 0.0, 0.0, 0.0, 0.0, f, f, width / 2.0, height / 2.0)
