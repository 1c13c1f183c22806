import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu() -> bool:
    try:
        from vins_fast_b200 import load
        return load().vf_device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # `-m gpu` on a box without a GPU must fail loudly instead of silently passing on nothing
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords and "gpu" not in (config.getoption("-m") or ""):
            it.add_marker(skip)


@pytest.fixture(scope="session")
def lib():
    from vins_fast_b200 import build
    build.build()
    from vins_fast_b200 import load
    return load()


@pytest.fixture(scope="session")
def oracle():
    from oracle import c_oracle
    c_oracle.build()
    return c_oracle


@pytest.fixture(scope="session")
def euroc_seq():
    from vins_fast_b200.synth import StereoSequence
    return StereoSequence(seed=0, width=752, height=480)


@pytest.fixture(scope="session")
def euroc_frames(euroc_seq):
    return [euroc_seq.frame(k) for k in range(24)]


def build_cpp_test(name: str, link_lib: bool = True) -> str:
    """g++ build of tests/cpp/<name>.cpp into tests/cpp/_build/ (rebuilt when a source or header is newer)."""
    import subprocess
    src = os.path.join(ROOT, "tests", "cpp", name + ".cpp")
    out_dir = os.path.join(ROOT, "tests", "cpp", "_build")
    os.makedirs(out_dir, exist_ok=True)
    exe = os.path.join(out_dir, name)
    deps = [src, os.path.join(ROOT, "include", "vins_fast_b200.h"), os.path.join(ROOT, "vins_fast_b200", "csrc", "introsort_emul.h")]
    inc = os.path.join(ROOT, "include", "vins_fast_b200")
    deps += [os.path.join(inc, f) for f in os.listdir(inc)]
    libdir = os.path.join(ROOT, "vins_fast_b200")
    if link_lib:
        from vins_fast_b200 import build
        deps.append(build.build())
    if not os.path.exists(exe) or any(os.path.getmtime(d) > os.path.getmtime(exe) for d in deps):
        cmd = ["g++", "-std=c++14", "-O2", "-Wall", "-I", os.path.join(ROOT, "include"), src, "-o", exe]
        if link_lib:
            cmd += ["-L", libdir, "-lvinsfast_b200", "-Wl,-rpath," + libdir]
        subprocess.check_call(cmd)
    return exe


def write_camera_yaml(path, cam, width: int, height: int):
    """camodocal calibration file (vins/config/euroc/cam0_mei.yaml layout) for an oracle.camera.Camera."""
    mei = cam.model == 0
    lines = ["%YAML:1.0", "---", f"model_type: {'MEI' if mei else 'PINHOLE'}", "camera_name: camera",
             f"image_width: {width}", f"image_height: {height}"]
    if mei:
        lines += ["mirror_parameters:", f"   xi: {cam.xi!r}"]
    lines += ["distortion_parameters:", f"   k1: {cam.k1!r}", f"   k2: {cam.k2!r}", f"   p1: {cam.p1!r}", f"   p2: {cam.p2!r}",
              "projection_parameters:"]
    names = ("gamma1", "gamma2", "u0", "v0") if mei else ("fx", "fy", "cx", "cy")
    lines += [f"   {n}: {v!r}" for n, v in zip(names, (cam.fx, cam.fy, cam.cx, cam.cy))]
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n")
    return path


def write_estimator_yaml(path, max_cnt: int, min_dist: int, flow_back: int, show_track: int = 0, width: int = 752, height: int = 480):
    """Estimator configuration in the layout of vins/config/euroc/euroc_stero.yaml (tracker keys + an opencv-matrix)."""
    text = f"""%YAML:1.0

#common parameters
imu: 0
num_of_cam: 2
cam0_calib: "cam0_mei.yaml"
cam1_calib: "cam1_mei.yaml"
image_width: {width}
image_height: {height}
body_T_cam0: !!opencv-matrix
   rows: 4
   cols: 4
   dt: d
   data: [0.0148655429818, -0.999880929698, 0.00414029679422, -0.0216401454975,
           0.999557249008, 0.0149672133247, 0.025715529948, -0.064676986768,
           0, 0, 0, 1]
max_cnt: {max_cnt}            # max feature number in feature tracking
min_dist: {min_dist}            # min distance between two features
flow_back: {flow_back}            # perform forward and backward optical flow to improve feature tracking accuracy
show_track: {show_track}
frontend_wait_key: 1
td: 0.5
"""
    with open(path, "w") as f:
        f.write(text)
    return path


def assert_frames_equal(a: dict, b: dict, ctx: str = ""):
    """Bit-exact comparison of two tracker outputs (oracle dict layout)."""
    for key in ("ids", "track_cnt", "uv", "un", "vel", "ids_right", "uv_right", "un_right", "vel_right"):
        x, y = np.asarray(a[key]), np.asarray(b[key])
        assert x.shape == y.shape, f"{ctx}: {key} shape {x.shape} != {y.shape}"
        if x.dtype.kind == "f":
            assert np.array_equal(x.view(np.uint32), y.view(np.uint32)), \
                f"{ctx}: {key} differs, max abs diff {np.abs(x.astype(np.float64) - y).max()}"
        else:
            assert np.array_equal(x, y), f"{ctx}: {key} differs"


FRAME_KEYS = ("ids", "track_cnt", "uv", "un", "vel", "ids_right", "uv_right", "un_right", "vel_right")


def frame_digest(frame: dict) -> bytes:
    """SHA-256 over the nine arrays of one TrackImage result (int32 / float32 bytes, fixed key order): the compact form
    in which the long cv2 sequences are kept under tests/golden/ (a checksum per frame instead of the arrays)."""
    import hashlib
    h = hashlib.sha256()
    for key in FRAME_KEYS:
        a = np.asarray(frame[key])
        a = a.astype(np.float32) if a.dtype.kind == "f" else a.astype(np.int32)
        h.update(np.ascontiguousarray(a).tobytes())
    return h.digest()


def first_divergence(frames_a, frames_b):
    """(frame index, key) of the first difference between two lists of tracker outputs, or None."""
    for k, (a, b) in enumerate(zip(frames_a, frames_b)):
        for key in FRAME_KEYS:
            x, y = np.asarray(a[key]), np.asarray(b[key])
            if x.shape != y.shape or x.tobytes() != y.astype(x.dtype).tobytes():
                return k, key
    return None


def write_euroc_dir(root, frames, jitter_ns=(0, 0), drop_right=(), extra_left=()):
    """An EuRoC-ASL-like directory: <root>/cam{0,1}/data.csv (timestamp [ns], filename) + binary PGM images.
    frames: [(t, left, right)].  jitter_ns: per-camera stamp offsets; drop_right: frame indices whose right image is lost;
    extra_left: indices after which an unpaired left image (stamp + 25 ms) is inserted.  Returns the expected pairs
    [(stamp0, left, right, stamp0 in ns)] under the 3 ms rule of Estimator::tFrontendProcess (estimator.cpp:103-141)."""
    import os
    expect = []
    lists = {0: [], 1: []}
    for cam in (0, 1):
        os.makedirs(os.path.join(root, f"cam{cam}", "data"), exist_ok=True)

    def put(cam, ns, img):
        name = f"{ns}.pgm"
        with open(os.path.join(root, f"cam{cam}", "data", name), "wb") as f:
            f.write(f"P5\n# synthetic\n{img.shape[1]} {img.shape[0]}\n255\n".encode() + np.ascontiguousarray(img).tobytes())
        lists[cam].append((ns, name))
    for k, (t, l, r) in enumerate(frames):
        ns = int(round(t * 1e9)) + 1_403_636_579_000_000_000
        put(0, ns + jitter_ns[0], l)
        if k not in drop_right:
            put(1, ns + jitter_ns[1], r)
            expect.append(((ns + jitter_ns[0]) * 1e-9, l, r, ns + jitter_ns[0]))
        if k in extra_left:
            put(0, ns + 25_000_000, 255 - l)
    for cam in (0, 1):
        with open(os.path.join(root, f"cam{cam}", "data.csv"), "w") as f:
            f.write("#timestamp [ns],filename\n" + "".join(f"{ns},{name}\n" for ns, name in lists[cam]))
    return expect
