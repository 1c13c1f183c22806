"""TEST INFRASTRUCTURE -- builds oracle/_ref/libvf_ref_geometry.so from the reference's own sources.

Only possible where /root/reference exists (this container; the GPU box uses the prebuilt file, which travels with the
snapshot).  Nothing of the reference is copied into the repository: the four function definitions are extracted into
oracle/_ref/ (git-ignored) at build time and compiled against the reference's vendored Eigen and its own headers.

    python -m oracle.build_ref
"""
from __future__ import annotations

import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("VF_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(HERE, "_ref")
SO = os.path.join(OUT, "libvf_ref_geometry.so")
FUNCS = ["void FeatureManager::TriangulatePts(", "void FeatureManager::TriangulateOnePoint(",
         "double FeatureManager::ReProjectionError(", "void FeatureManager::OutliersRejection("]


def extract(src: str) -> str:
    """The definitions of FUNCS: from the signature line to the closing brace in column 0."""
    lines = src.splitlines()
    out = ["// GENERATED from the reference's vins/estimator/feature_manager.cpp by oracle/build_ref.py -- do not commit",
           "namespace estimator {"]
    for sig in FUNCS:
        start = next(i for i, l in enumerate(lines) if l.startswith(sig))
        end = next(i for i in range(start, len(lines)) if lines[i].startswith("}"))
        out += lines[start:end + 1] + [""]
    out.append("}  // namespace estimator")
    return "\n".join(out) + "\n"


def available() -> bool:
    return os.path.exists(os.path.join(REF, "vins", "estimator", "feature_manager.cpp"))


def build(force: bool = False) -> str | None:
    """Returns the library path, or None when neither the reference nor a prebuilt library is there."""
    if not available():
        return SO if os.path.exists(SO) else None
    src = os.path.join(REF, "vins", "estimator", "feature_manager.cpp")
    wrapper = os.path.join(HERE, "ref_geometry.cpp")
    if not force and os.path.exists(SO) and os.path.getmtime(SO) >= max(os.path.getmtime(wrapper), os.path.getmtime(__file__)):
        return SO
    os.makedirs(OUT, exist_ok=True)
    with open(src) as f:
        inc = extract(f.read())
    with open(os.path.join(OUT, "feature_manager_extract.inc"), "w") as f:
        f.write(inc)
    cmd = ["g++", "-O2", "-std=c++14", "-fPIC", "-shared", "-ffp-contract=off", "-w",
           "-I", os.path.join(REF, "vins"), "-I", os.path.join(REF, "vins", "thirdparty", "eigen"), "-I", OUT,
           wrapper, "-o", SO]
    subprocess.check_call(cmd)
    return SO


if __name__ == "__main__":
    print(build(force=True))
