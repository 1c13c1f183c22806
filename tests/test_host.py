"""CPU suite for everything that is not a kernel: the C-ABI library loads and exports every symbol the header declares,
argument validation and the no-GPU failure mode, the OpenCV-dialect YAML reader, the libstdc++ introsort replay, the
header-only C++ FeatureTracker / common::Setting host layer, the stream partition used for multi-GPU runs (gloo,
world_size 2), and the synthetic input generator.  No compute call needs a GPU here."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, build_cpp_test, write_camera_yaml, write_estimator_yaml
from oracle.camera import EUROC_CAM0, EUROC_CAM1, pinhole_for


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "vins_fast_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vf_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(lib):
    from vins_fast_b200 import _lib
    declared = _header_symbols()
    assert len(declared) >= 35
    assert sorted(_lib.EXPORTS) == declared, set(_lib.EXPORTS) ^ set(declared)
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in include/vins_fast_b200.h but not exported"


def test_library_is_built_for_sm_100a():
    so = os.path.join(ROOT, "vins_fast_b200", "libvinsfast_b200.so")
    out = subprocess.run(["cuobjdump", "--list-elf", so], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in out.stdout


def test_struct_layouts_match_the_header(lib):
    from vins_fast_b200 import _lib
    assert C.sizeof(_lib.VfFeature) == 60 and _lib.FEATURE_DTYPE.itemsize == 60
    assert C.sizeof(_lib.VfCamera) == 88
    cfg = _lib.VfConfig()
    lib.vf_config_default(C.byref(cfg))
    # euroc_stero.yaml:44-51 defaults and the reference's literal call arguments (feature_tracker.cpp:42,90)
    assert (cfg.width, cfg.height, cfg.max_cnt, cfg.min_dist, cfg.flow_back) == (752, 480, 200, 15, 1)
    assert (cfg.lk_max_level, cfg.lk_win, cfg.quality_level, cfg.device, cfg.use_cuda_graph) == (3, 21, 0.01, -1, 1)


def test_argument_validation_needs_no_gpu(lib):
    from vins_fast_b200 import _lib
    assert lib.vf_tracker_create(None, None) == _lib.VF_ERR_INVALID_ARG
    h = C.c_void_p()
    assert lib.vf_tracker_create(None, C.byref(h)) == _lib.VF_ERR_INVALID_ARG and b"null" in lib.vf_last_global_error()
    cfg = _lib.VfConfig()
    lib.vf_config_default(C.byref(cfg))
    for field, bad in (("lk_win", 14), ("lk_win", 33), ("lk_win", 3), ("max_cnt", 0), ("min_dist", -1), ("width", 10), ("quality_level", 0.0), ("lk_max_level", 9)):
        c2 = _lib.VfConfig.from_buffer_copy(cfg)
        setattr(c2, field, bad)
        assert lib.vf_tracker_create(C.byref(c2), C.byref(h)) == _lib.VF_ERR_INVALID_ARG, field
        assert not h.value
    c2 = _lib.VfConfig.from_buffer_copy(cfg)
    c2.cam[1].model = 7
    assert lib.vf_tracker_create(C.byref(c2), C.byref(h)) == _lib.VF_ERR_INVALID_ARG
    assert lib.vf_batch_create(C.byref(cfg), 0, C.byref(h)) == _lib.VF_ERR_INVALID_ARG
    assert lib.vf_op_build_pyramid(None, 10, 10, 3, -1, None, None, None, None) == _lib.VF_ERR_INVALID_ARG
    assert lib.vf_op_lk(None, None, 10, 10, None, None, None, 0, 3, 0, -1, None) == _lib.VF_ERR_INVALID_ARG
    assert lib.vf_op_undistort(None, None, 0, -1, None) == _lib.VF_ERR_INVALID_ARG
    assert lib.vf_tracker_track(None, 0.0, None, 0, None, 0, None, None) == _lib.VF_ERR_INVALID_ARG
    assert lib.vf_batch_size(None) == 0 and lib.vf_batch_stage_count() >= 10
    lib.vf_tracker_destroy(None)
    lib.vf_batch_destroy(None)


def test_no_gpu_fails_loudly_never_falls_back(lib):
    """The product path must not route through any CPU implementation: without a device every compute entry point
    returns VF_ERR_CUDA with a message."""
    from vins_fast_b200 import _lib
    if lib.vf_device_count() > 0:
        pytest.skip("a CUDA device is present")
    cfg = _lib.VfConfig()
    lib.vf_config_default(C.byref(cfg))
    h = C.c_void_p()
    assert lib.vf_tracker_create(C.byref(cfg), C.byref(h)) == _lib.VF_ERR_CUDA
    assert b"no CUDA device" in lib.vf_last_global_error() and b"no CPU fallback" in lib.vf_last_global_error()
    img = np.zeros((64, 64), np.uint8)
    out = np.zeros(2 * 64 * 64, np.uint8)
    ws, hs, n = (C.c_int32 * 8)(), (C.c_int32 * 8)(), C.c_int32()
    assert lib.vf_op_build_pyramid(img.ctypes.data, 64, 64, 1, -1, out.ctypes.data, ws, hs, C.byref(n)) == _lib.VF_ERR_CUDA
    import vins_fast_b200 as v
    with pytest.raises(v.VfError) as e:
        v.FeatureTracker(v.make_config(752, 480, 150, 30)).track(0.0, np.zeros((480, 752), np.uint8), np.zeros((480, 752), np.uint8))
    assert e.value.status == _lib.VF_ERR_CUDA


def test_product_package_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under vins_fast_b200/ or include/ may import, link or execute it."""
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|oracle/|vf_oracle|libvf_oracle", re.M)
    for base in ("vins_fast_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                    text = open(os.path.join(dirpath, f)).read()
                    hits = [m.group(0) for m in pat.finditer(text) if "oracle/vf_oracle.c" not in text[max(0, m.start() - 60):m.end() + 20]]
                    assert not hits, f"{base}/{f} references the oracle: {hits}"


# ---- YAML (common::Setting / camodocal readers) -------------------------------------------------------------------
def test_camera_yaml_roundtrip(lib, tmp_path):
    import vins_fast_b200 as v
    from vins_fast_b200 import _lib
    for k, cam in enumerate((EUROC_CAM0, EUROC_CAM1, pinhole_for(1280, 720))):
        p = write_camera_yaml(tmp_path / f"cam{k}.yaml", cam, 752, 480)
        c = v.load_camera_yaml(str(p))
        assert (c.model, c.image_width, c.image_height) == (cam.model, 752, 480)
        for f in ("k1", "k2", "p1", "p2", "fx", "fy", "cx", "cy"):
            assert getattr(c, f) == getattr(cam, f), f
        assert c.xi == (cam.xi if cam.model == 0 else 0.0)
    bad = tmp_path / "bad.yaml"
    bad.write_text("%YAML:1.0\n---\nmodel_type: KANNALA_BRANDT\n")
    with pytest.raises(v.VfError) as e:
        v.load_camera_yaml(str(bad))
    assert e.value.status == _lib.VF_ERR_PARSE and "unsupported camera model_type" in e.value.message
    with pytest.raises(v.VfError) as e:
        v.load_camera_yaml(str(tmp_path / "missing.yaml"))
    assert e.value.status == _lib.VF_ERR_IO
    lower = tmp_path / "lower.yaml"
    lower.write_text(open(write_camera_yaml(tmp_path / "c.yaml", EUROC_CAM0, 752, 480)).read().replace("MEI", "mei"))
    assert v.load_camera_yaml(str(lower)).model == 0                     # CameraFactory.cc:96-136: case-insensitive


def test_estimator_yaml(lib, tmp_path):
    import vins_fast_b200 as v
    from vins_fast_b200 import _lib
    write_camera_yaml(tmp_path / "cam0_mei.yaml", EUROC_CAM0, 752, 480)
    write_camera_yaml(tmp_path / "cam1_mei.yaml", EUROC_CAM1, 752, 480)
    est = write_estimator_yaml(tmp_path / "est.yaml", 150, 30, 1, show_track=0)
    cfg = v.load_config_yaml(str(est))
    assert (cfg.width, cfg.height, cfg.max_cnt, cfg.min_dist, cfg.flow_back, cfg.lk_max_level) == (752, 480, 150, 30, 1, 3)
    assert cfg.cam[0].xi == EUROC_CAM0.xi and cfg.cam[1].fx == EUROC_CAM1.fx       # cam0_calib / cam1_calib next to the file
    assert (cfg.lk_win, cfg.quality_level) == (21, 0.01)                            # the reference's literals when the keys are absent
    keyed = tmp_path / "est_keys.yaml"                                              # north star: pyramid levels and window size are parameters
    keyed.write_text(est.read_text() + "lk_max_level: 4\nlk_win: 15\nquality_level: 0.02\n")
    cfg2 = v.load_config_yaml(str(keyed))
    assert (cfg2.lk_max_level, cfg2.lk_win, cfg2.quality_level) == (4, 15, 0.02)
    buf = C.create_string_buffer(64)
    assert lib.vf_yaml_get(str(est).encode(), b"max_cnt", buf, 64) == _lib.VF_OK and buf.value == b"150"
    assert lib.vf_yaml_get(str(est).encode(), b"cam0_calib", buf, 64) == _lib.VF_OK and buf.value == b"cam0_mei.yaml"
    assert lib.vf_yaml_get(str(est).encode(), b"nope", buf, 64) == _lib.VF_ERR_PARSE
    assert lib.vf_yaml_get(str(est).encode(), b"max_cnt", buf, 2) == _lib.VF_ERR_INVALID_ARG
    missing = tmp_path / "nokeys.yaml"
    missing.write_text("%YAML:1.0\nimu: 0\n")
    with pytest.raises(v.VfError) as e:
        v.load_config_yaml(str(missing))
    assert e.value.status == _lib.VF_ERR_PARSE
    # a calibration file the YAML lists but nobody can open is an error, as in CameraFactory::generateCameraFromYamlFile
    # (the shipped configs carry absolute /home/weihao/... paths): never a silent default pinhole
    (tmp_path / "lost").mkdir()
    lost = write_estimator_yaml(tmp_path / "lost" / "est.yaml", 150, 30, 1)
    with pytest.raises(v.VfError) as e:
        v.load_config_yaml(str(lost))
    assert e.value.status == _lib.VF_ERR_IO and "cam0_mei.yaml" in e.value.message


# ---- libstdc++ introsort replay (host build of the device code) ---------------------------------------------------
def test_introsort_emulation_equals_std_sort():
    exe = build_cpp_test("test_introsort", link_lib=False)
    res = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr


# ---- C++ drop-in class: host layer ------------------------------------------------------------------------------------
def test_cpp_feature_tracker_host_layer(lib, tmp_path):
    exe = build_cpp_test("test_feature_tracker")
    write_camera_yaml(tmp_path / "cam0_mei.yaml", EUROC_CAM0, 752, 480)
    write_camera_yaml(tmp_path / "cam1_mei.yaml", EUROC_CAM1, 752, 480)
    est = write_estimator_yaml(tmp_path / "est.yaml", 150, 30, 1, show_track=1)
    res = subprocess.run([exe, "host", str(est), str(tmp_path / "cam0_mei.yaml"), str(tmp_path / "cam1_mei.yaml")],
                         capture_output=True, text=True, timeout=120)
    assert res.returncode == 0 and "host ok" in res.stdout, res.stdout + res.stderr


def test_cpp_header_compiles_as_c_and_cpp11(tmp_path):
    """include/vins_fast_b200.h is plain C; the C++ layer builds with the reference's -std=c++14 and with c++11."""
    c = tmp_path / "abi.c"
    c.write_text('#include "vins_fast_b200.h"\nint main(void) { vf_config c; vf_config_default(&c); return sizeof(vf_feature) == 60 ? 0 : 1; }\n')
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), "-c", str(c), "-o", str(tmp_path / "abi.o")])
    cpp = tmp_path / "hdr.cpp"
    cpp.write_text('#include "vins_fast_b200/feature_tracker.hpp"\nint main() { return 0; }\n')
    for std in ("c++11", "c++14", "c++17"):
        subprocess.check_call(["g++", f"-std={std}", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", str(cpp), "-o", str(tmp_path / "hdr.o")])


# ---- multi-GPU partition: independent streams, no data-path collective ------------------------
def test_stream_partition():
    from vins_fast_b200.shard import streams_for_rank
    for S in (1, 7, 64):
        for G in (1, 2, 4, 8):
            parts = [streams_for_rank(S, r, G) for r in range(G)]
            assert sorted(sum(parts, [])) == list(range(S))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
            for r, p in enumerate(parts):
                assert all(s % G == r for s in p)                         # stream s -> GPU s mod G


_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import numpy as np
import torch.distributed as dist
from vins_fast_b200.shard import streams_for_rank, gather_stream_results, max_over_ranks
from vins_fast_b200.synth import StereoSequence
from oracle import c_oracle
from oracle.camera import pinhole_for
dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
rank, world = dist.get_rank(), dist.get_world_size()
S, w, h = 3, 160, 120
mine = streams_for_rank(S, rank, world)
cam = pinhole_for(w, h)
res = {{}}
for s in mine:                       # the CPU oracle stands in for the GPU tracker: the partition logic is what is tested
    seq = StereoSequence(100 + s, w, h)
    tr = c_oracle.CFeatureTracker(cam, cam, w, h, 30, 10, 1, 2)
    ids = []
    for k in range(3):
        ids.append(tr.track_image(*seq.frame(k))["ids"].copy())
    res[s] = np.concatenate(ids)
allres = gather_stream_results(res, S)
t = max_over_ranks(float(rank + 1))
if rank == 0:
    assert sorted(allres) == list(range(S))
    assert t == float(world)
    np.savez({out!r}, **{{f"s{{s}}": v for s, v in allres.items()}})
dist.destroy_process_group()
"""


def test_two_rank_gloo_shards_streams_without_collectives_on_the_data_path(oracle, tmp_path):
    pytest.importorskip("torch")
    out = str(tmp_path / "gathered.npz")
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT, out=out))
    port = 29500 + os.getpid() % 2000
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    for p in procs:
        o, _ = p.communicate(timeout=300)
        assert p.returncode == 0, o
    got = np.load(out)
    # same streams run in one process give the same answer: shards are independent
    from vins_fast_b200.synth import StereoSequence
    cam = pinhole_for(160, 120)
    for s in range(3):
        seq = StereoSequence(100 + s, 160, 120)
        tr = oracle.CFeatureTracker(cam, cam, 160, 120, 30, 10, 1, 2)
        ids = np.concatenate([tr.track_image(*seq.frame(k))["ids"] for k in range(3)])
        assert np.array_equal(got[f"s{s}"], ids)


# ---- synthetic input ----------------------------------------------------------------------------------------------------
def test_synthetic_sequence_is_deterministic_and_textured():
    import hashlib
    from vins_fast_b200.synth import StereoSequence
    a, b = StereoSequence(0, 752, 480), StereoSequence(0, 752, 480)
    t, l, r = a.frame(5)
    t2, l2, r2 = b.frame(5)
    assert t == 0.25 and l.tobytes() == l2.tobytes() and r.tobytes() == r2.tobytes()
    assert l.shape == (480, 752) and l.dtype == np.uint8 and l.flags.c_contiguous
    assert l.std() > 20 and not np.array_equal(l, r)
    g = np.load(os.path.join(ROOT, "tests", "golden", "sequence_euroc_mei_cv2_4_13.npz"))
    t0, l0, r0 = a.frame(0)
    sha = hashlib.sha256(l0.tobytes()).hexdigest() + hashlib.sha256(r0.tobytes()).hexdigest()
    assert bytes.fromhex(sha) == g["in_sha_0"].tobytes()


def test_cpp_header_compiles_against_the_reference_eigen(tmp_path):
    """With the reference's vendored Eigen 3.3.7 on the include path compat.hpp uses the real Eigen::Matrix (this
    container only: /root/reference does not exist on the GPU box)."""
    eigen = "/root/reference/vins/thirdparty/eigen"
    if not os.path.isdir(os.path.join(eigen, "Eigen")):
        pytest.skip("reference tree not present")
    cpp = tmp_path / "eig.cpp"
    cpp.write_text('#define VINS_FAST_B200_THROW 1\n#include "vins_fast_b200/feature_tracker.hpp"\n'
                   '#ifndef VF_HAVE_EIGEN\n#error real Eigen was not picked up\n#endif\n'
                   'int main() { Eigen::Matrix<double, 7, 1> v; v << 1, 2, 3, 4, 5, 6, 7; estimator::FeatureTracker ft(150, 30, 1);\n'
                   '  return v(6) == 7.0 && ft.MIN_DIST == 30 ? 0 : 1; }\n')
    exe = tmp_path / "eig"
    libdir = os.path.join(ROOT, "vins_fast_b200")
    subprocess.check_call(["g++", "-std=c++14", "-w", "-I", os.path.join(ROOT, "include"), "-I", eigen, str(cpp), "-o", str(exe),
                           "-L", libdir, "-lvinsfast_b200", "-Wl,-rpath," + libdir])
    assert subprocess.run([str(exe)]).returncode == 0


def test_feature_frames_feed_the_reference_consumer_types(tmp_path):
    """featureFrames from the host layer's ToFeatureFrame go through the reference's unchanged
    ObservedFrameFeature / IDWithObservedFeatures (compiled from the reference tree, this container only) with the calls
    FeatureManager::AddFeatureAndCheckParallax makes (feature_manager.cpp:25-62); the bookkeeping it derives must match
    what the golden sequence implies."""
    ref = "/root/reference/vins"
    if not os.path.isfile(os.path.join(ref, "estimator", "id_features.hpp")):
        pytest.skip("reference tree not present")
    from vins_fast_b200 import FEATURE_DTYPE
    g = np.load(os.path.join(ROOT, "tests", "golden", "sequence_euroc_mei_cv2_4_13.npz"))
    n_frames = len([k for k in g.keys() if k.endswith("_ids")])
    seen, want = {}, []
    with open(tmp_path / "records.bin", "wb") as f:
        for k in range(n_frames):
            ids = g[f"f{k}_ids"]
            rec = np.zeros(len(ids), FEATURE_DTYPE)
            rec["id"], rec["track_cnt"] = ids, g[f"f{k}_track_cnt"]
            rec["u"], rec["v"] = g[f"f{k}_uv"].T
            rec["x"], rec["y"] = g[f"f{k}_un"].T
            rec["vx"], rec["vy"] = g[f"f{k}_vel"].T
            pos = {int(i): j for j, i in enumerate(ids)}
            rows = [pos[int(i)] for i in g[f"f{k}_ids_right"]]
            rec["has_right"][rows] = 1
            rec["ur"][rows], rec["vr"][rows] = g[f"f{k}_uv_right"].T
            rec["xr"][rows], rec["yr"][rows] = g[f"f{k}_un_right"].T
            rec["vxr"][rows], rec["vyr"][rows] = g[f"f{k}_vel_right"].T
            f.write(np.int32(len(rec)).tobytes())
            f.write(rec.tobytes())
            last = sum(1 for i in ids if int(i) in seen)
            for i in ids:
                seen[int(i)] = seen.get(int(i), 0) + 1
            long_ = sum(1 for i in ids if seen[int(i)] >= 4)
            want.append((last, len(ids) - last, long_, len(rows)))
    exe = tmp_path / "consumer"
    subprocess.check_call(["g++", "-std=c++14", "-w", "-I", os.path.join(ROOT, "include"), "-I", ref,
                           "-I", os.path.join(ref, "thirdparty", "eigen"),
                           os.path.join(ROOT, "tests", "cpp", "test_consumer_types.cpp"), "-o", str(exe),
                           "-L", os.path.join(ROOT, "vins_fast_b200"), "-lvinsfast_b200",
                           "-Wl,-rpath," + os.path.join(ROOT, "vins_fast_b200")])
    out = subprocess.run([str(exe), str(tmp_path / "records.bin"), str(n_frames)], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    lines = [l.split() for l in out.stdout.strip().splitlines()]
    assert len(lines) == n_frames
    for k, l in enumerate(lines):
        assert tuple(int(v) for v in l[:4]) == want[k], (k, l, want[k])
        rec_sum = sum(float(np.float64(g[f"f{k}_{n}"]).sum()) for n in ("un", "uv", "vel"))
        rec_sum += sum(float(np.float64(g[f"f{k}_{n}_right"][:, 0]).sum()) for n in ("un", "uv", "vel"))
        assert abs(float(l[4]) - rec_sum) < 1e-6 * max(1.0, abs(rec_sum))
    assert want[0][1] == len(g["f0_ids"]) and any(w[2] > 0 for w in want)      # all new on frame 0, long tracks later


def test_draw_tracker_is_pixel_identical_to_opencv(tmp_path):
    """DrawTracker (feature_tracker.cpp:267-299).  The C++ renderer (include/vins_fast_b200/draw.hpp) on
    random tracker-like scenes against the same cv2 calls the reference makes: hconcat, cvtColor(GRAY2RGB),
    circle(r 2, thickness 2), arrowedLine(thickness 1, LINE_8, tipLength 0.2)."""
    cv2 = pytest.importorskip("cv2")
    from conftest import build_cpp_test
    exe = build_cpp_test("test_draw", link_lib=False)
    rng = np.random.default_rng(17)
    for case, (w, h, n) in enumerate(((752, 480, 150), (320, 240, 60), (97, 64, 25))):
        left = rng.integers(0, 256, (h, w), dtype=np.uint8)
        right = rng.integers(0, 256, (h, w), dtype=np.uint8)
        # in-border points (InBorder, feature_tracker.cpp:20-25), half-pixel cases included
        kps = np.stack([rng.uniform(1, w - 2, n), rng.uniform(1, h - 2, n)], 1).astype(np.float32)
        kps[:6] = np.floor(kps[:6]) + 0.5
        cnt = rng.integers(1, 40, n).astype(np.int32)
        cnt[:4] = (1, 10, 20, 21)
        nr = n - n // 5
        rk = np.stack([rng.uniform(1, w - 2, nr), rng.uniform(1, h - 2, nr)], 1).astype(np.float32)
        step = rng.uniform(-9, 9, (n, 2)) * rng.choice([0.0, 0.3, 1.0, 4.0, 12.0], (n, 1))   # long arrows: tips leave the image
        prev = np.clip(kps + step, 1, [w - 2, h - 2]).astype(np.float32)
        has_prev = (rng.uniform(0, 1, n) < 0.85).astype(np.uint8)
        scene = tmp_path / f"scene{case}.bin"
        with open(scene, "wb") as f:
            f.write(np.array([w, h, n, nr], np.int32).tobytes() + left.tobytes() + right.tobytes() + kps.tobytes() + cnt.tobytes()
                    + rk.tobytes() + prev.tobytes() + has_prev.tobytes())
        out = tmp_path / f"out{case}.rgb"
        subprocess.check_call([exe, str(scene), str(out)])
        got = np.fromfile(out, np.uint8).reshape(h, 2 * w, 3)
        # the reference's calls, verbatim (Point2f arguments are rounded by cv::Point's converting constructor = cvRound)
        rnd = lambda p: (int(np.rint(np.float32(p[0]))), int(np.rint(np.float32(p[1]))))
        img = cv2.cvtColor(cv2.hconcat([left, right]), cv2.COLOR_GRAY2RGB)
        for j in range(n):
            ln = min(1.0, 1.0 * int(cnt[j]) / 20)
            cv2.circle(img, rnd(kps[j]), 2, (255 * (1 - ln), 0, 255 * ln), 2)
        for j in range(nr):
            cv2.circle(img, rnd((np.float32(rk[j, 0]) + np.float32(w), rk[j, 1])), 2, (0, 255, 0), 2)
        for i in range(n):
            if has_prev[i]:
                cv2.arrowedLine(img, rnd(kps[i]), rnd(prev[i]), (0, 255, 0), 1, 8, 0, 0.2)
        assert np.array_equal(got, img), f"case {case}: {int((got != img).any(axis=2).sum())} pixels differ"


def test_input_side_pgm_euroc_listing_and_stereo_sync(tmp_path):
    """SURVEY 8(f) rank 2 on the host: EuRoC data.csv listings + raw PGM images -> StereoSync (the 3 ms pairing rule of
    Estimator::tFrontendProcess, estimator.cpp:103-141) -> synchronised pairs; GetImageFromROSMsg's mono8 / 8UC1 copy
    (vins/common/utils.hpp:17-37)."""
    from conftest import build_cpp_test, write_euroc_dir
    from vins_fast_b200.synth import StereoSequence
    exe = build_cpp_test("test_image_source")
    seq = StereoSequence(4, 160, 120)
    frames = [seq.frame(k) for k in range(9)]
    expect = write_euroc_dir(str(tmp_path), frames, jitter_ns=(0, 2_000_000), drop_right=(3,), extra_left=(6,))
    res = subprocess.run([exe, "host", str(tmp_path)], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0 and "host ok" in res.stdout, res.stdout + res.stderr
    pairs = [ln.split() for ln in res.stdout.splitlines() if ln.startswith("pair")]
    assert len(pairs) == len(expect) == 8
    for got, (t, l, r, _) in zip(pairs, expect):
        assert abs(float(got[1]) - t) < 1e-6 and int(got[2]) == int(l.sum()) and int(got[3]) == int(r.sum())
    assert "thrown 2 0" in res.stdout          # the left image whose partner was lost + the unpaired extra one


def test_cpu_affinity_helper_degrades_without_nvml():
    """bench.py runs every rank on the CPUs next to its GPU (vins_fast_b200/affinity.py); without NVML / a GPU the helper
    reports why it did nothing and leaves the process where it is."""
    import os
    from vins_fast_b200.affinity import bind_to_gpu, gpu_cpu_set
    before = os.sched_getaffinity(0)
    info = bind_to_gpu(0)
    assert set(info) >= {"bound"} and (info["bound"] or "why" in info)
    if not gpu_cpu_set(0):
        assert info["bound"] is False and os.sched_getaffinity(0) == before
