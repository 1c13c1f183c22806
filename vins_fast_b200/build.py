"""Builds libvinsfast_b200.so (CUDA kernels + C ABI; the C++ FeatureTracker of include/vins_fast_b200/ is header-only over that ABI) in-tree for sm_100a with nvcc.

    python -m vins_fast_b200.build [--force] [--verbose]

The .so is git-ignored but travels to the GPU box with the repository snapshot.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "libvinsfast_b200.so")

SOURCES = ["vf_api.cu", "vf_pyramid.cu", "vf_pyramid_tma.cu", "vf_lk.cu", "vf_lk_warp.cu", "vf_lk_generic.cu", "vf_corners.cu", "vf_select.cu", "vf_geometry.cu", "vf_yaml.cpp"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false",            # parity: no FMA contraction anywhere (critical expressions also use _rn intrinsics)
    "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=default", "-cudart", "static", "-shared",
]


STAMP = LIB + ".hash"      # sha256 of every source + the flags the library was built from (travels with the .so)


def _deps():
    out = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC))
    out.append(os.path.join(ROOT, "include", "vins_fast_b200.h"))
    return out


def source_hash() -> str:
    """Identity of a build: the bytes of every file under csrc/, the C header and the nvcc flags.  File times mean nothing on
    a fresh snapshot of the repository, so the decision to rebuild is taken on this hash, stored beside the library."""
    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS + SOURCES).encode())
    for d in _deps():
        h.update(os.path.basename(d).encode() + b"\0")
        with open(d, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def needs_build() -> bool:
    if not os.path.exists(LIB) or not os.path.exists(STAMP):
        return True
    with open(STAMP) as f:
        return f.read().strip() != source_hash()


def build_timing() -> str:
    """Measurement aid: a second library with the LK phase counters compiled in (-DVF_LKW_TIMING), selected with VF_LIB=<path>."""
    out = os.path.join(HERE, "libvinsfast_b200_timing.so")
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run([nvcc] + NVCC_FLAGS + ["-DVF_LKW_TIMING", "-I", os.path.join(ROOT, "include"), "-o", out] + srcs, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building the timing library")
    return out


def build_variant(name: str, defines: list) -> str:
    """Measurement aid: libvinsfast_b200_<name>.so built with extra -D flags, selected with VF_LIB=<path> (A/B runs on the GPU box)."""
    out = os.path.join(HERE, f"libvinsfast_b200_{name}.so")
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run([nvcc] + NVCC_FLAGS + list(defines) + ["-I", os.path.join(ROOT, "include"), "-o", out] + srcs, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError(f"nvcc failed building the {name} variant")
    return out


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-I", os.path.join(ROOT, "include"), "-o", LIB] + srcs
    if verbose:
        print(" ".join(cmd))
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libvinsfast_b200.so")
    if verbose:
        print(res.stdout + res.stderr)
    with open(STAMP, "w") as f:
        f.write(source_hash() + "\n")
    return LIB


if __name__ == "__main__":
    if "--variant" in sys.argv:      # python -m vins_fast_b200.build --variant NAME -DX=1 -DY=2
        i = sys.argv.index("--variant")
        print(build_variant(sys.argv[i + 1], [a for a in sys.argv[i + 2:] if a.startswith("-D")]))
        sys.exit(0)
    if "--timing" in sys.argv:
        print(build_timing())
        sys.exit(0)
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
