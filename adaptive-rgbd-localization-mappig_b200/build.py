"""Builds liborbfront_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

  python adaptive-rgbd-localization-mappig_b200/build.py [--force]

-fmad=false: the RANSAC / descriptor arithmetic must not contract a*b+c into FMA; the
integer kernels are unaffected.  -lineinfo keeps ncu's source page mapped to these files.
"""
import os
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
OUT = HERE / "liborbfront_b200.so"
SOURCES = ["context.cu", "pyramid.cu", "fast.cu", "quadtree.cu", "describe.cu", "match.cu", "ransac.cu", "kfdb.cu", "adaptive.cu", "projection.cu", "c_abi.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-fmad=false",
         "-Xcompiler", "-fPIC,-ffp-contract=off,-Wall,-Wno-unused-function", "-Xptxas", "-v", "--threads", "0"]


def needs_build():
    if not OUT.exists():
        return True
    t = OUT.stat().st_mtime
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.h")) + list(CSRC.glob("*.inc")) + [HERE.parent / "include" / "orbfront.h", Path(__file__)]
    return any(d.stat().st_mtime > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return str(OUT)
    cmd = [NVCC] + FLAGS + ["-shared", "-o", str(OUT)] + [str(CSRC / s) for s in SOURCES]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log = r.stdout + r.stderr
    (HERE / "build.log").write_text(log)
    if r.returncode != 0:
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed")
    if verbose:
        print(log)
    return str(OUT)


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
