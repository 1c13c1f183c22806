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
