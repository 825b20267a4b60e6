"""Builds liborbfront_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

  python adaptive-rgbd-localization-mappig_b200/build.py [--force]

-fmad=false: the RANSAC / descriptor arithmetic must not contract a*b+c into FMA; the
integer kernels are unaffected.  -lineinfo keeps ncu's source page mapped to these files.
"""
import fcntl
import hashlib
import os
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
OUT = HERE / "liborbfront_b200.so"
SOURCES = ["context.cu", "pyramid.cu", "fast.cu", "quadtree.cu", "describe.cu", "match.cu", "ransac.cu", "kfdb.cu", "comm.cu", "adaptive.cu", "projection.cu", "c_abi.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-fmad=false",
         "-Xcompiler", "-fPIC,-ffp-contract=off,-Wall,-Wno-unused-function", "-Xptxas", "-v", "--threads", "0"]


STAMP = HERE / "liborbfront_b200.so.stamp"          # sha256 of every input of the last build: travels with the .so to the GPU box


def _deps():
    return sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.h")) + list(CSRC.glob("*.inc"))) + [HERE.parent / "include" / "orbfront.h", Path(__file__)]


def source_hash():
    h = hashlib.sha256()
    h.update(" ".join(FLAGS).encode())
    for d in _deps():
        h.update(d.name.encode()); h.update(d.read_bytes())
    return h.hexdigest()


def needs_build():
    """Content-based, not mtime-based: a snapshot copied to another machine must not look stale (N ranks started by torchrun
    would otherwise all recompile, into the same file)."""
    if not OUT.exists() or not STAMP.exists():
        return True
    return STAMP.read_text().strip() != source_hash()


def build(force=False, verbose=False):
    if not force and not needs_build():
        return str(OUT)
    # one builder at a time (ranks of one node share the tree); the others wait for the lock, find the stamp current and return
    with open(HERE / ".build.lock", "w") as lock:
        try:
            fcntl.flock(lock, fcntl.LOCK_EX)
        except OSError:
            pass                                      # a filesystem without flock: the atomic rename below still keeps the file whole
        if not force and not needs_build():
            return str(OUT)
        want = source_hash()
        tmp = HERE / f"liborbfront_b200.so.tmp.{os.getpid()}"
        cmd = [NVCC] + FLAGS + ["-shared", "-o", str(tmp)] + [str(CSRC / s) for s in SOURCES] + ["-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log = r.stdout + r.stderr
        (HERE / "build.log").write_text(log)
        if r.returncode != 0:
            if tmp.exists():
                tmp.unlink()
            sys.stderr.write(log)
            raise RuntimeError("nvcc failed")
        os.replace(tmp, OUT)                          # atomic: a concurrent dlopen sees the old or the new file, never half of one
        STAMP.write_text(want + "\n")
        if verbose:
            print(log)
    return str(OUT)


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
