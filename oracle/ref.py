"""ctypes binding of oracle/_ref/liborb_ref.so — TEST INFRASTRUCTURE ONLY.

liborb_ref.so is the reference's own Features/orbextractor.cpp, compiled verbatim from /root/reference (oracle/Makefile, target
_ref) against the OpenCV stand-in under oracle/ref_shim/.  Only tests/, bench.py's CPU legs and __graft_entry__.build() touch it.
It exists where the reference checkout exists (the authoring container); the prebuilt file travels to the GPU box with the
snapshot (oracle/_ref/ is git-ignored, not gpurun-ignored).  available() says whether it can be used.
"""
import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

from .oracle import CAND_DT, KEYPOINT_DT, level_sizes, _p

_DIR = Path(__file__).resolve().parent
SO = _DIR / "_ref" / "liborb_ref.so"
REFERENCE = Path(os.environ.get("ORB_REFERENCE_DIR", "/root/reference"))
_lib = None


def build():
    """Compile from the reference checkout when it is present; otherwise keep whatever prebuilt file is there."""
    if (REFERENCE / "Features" / "orbextractor.cpp").exists():
        subprocess.run(["make", "-C", str(_DIR), "-s", "_ref", f"REF={REFERENCE}"], check=True)
    return SO.exists()


def available():
    return SO.exists() or build()


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("oracle/_ref/liborb_ref.so is not built and the reference checkout is absent")
        _lib = C.CDLL(str(SO))
    return _lib


def extract(img, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7, want_pyramid=False):
    """ORBextractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST)(img, noArray(), keypoints, descriptors)."""
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    cap = 2 * nfeatures + 64 * nlevels
    kps = np.zeros(cap, KEYPOINT_DT); desc = np.zeros((cap, 32), np.uint8); n = C.c_int(0)
    pyr = None
    if want_pyramid:
        ws, hs = level_sizes(w, h, scale_factor, nlevels)
        pyr = np.zeros(int((ws.astype(np.int64) * hs).sum()), np.uint8)
    rc = lib().ref_orb_extract(_p(img), w, h, w, nfeatures, C.c_float(scale_factor), nlevels, ini_th, min_th, _p(kps), _p(desc), cap,
                               C.byref(n), _p(pyr), None)
    if rc:
        raise RuntimeError(f"ref_orb_extract rc={rc}")
    if want_pyramid:
        return kps[:n.value].copy(), desc[:n.value].copy(), pyr
    return kps[:n.value].copy(), desc[:n.value].copy()


def distribute(cands, min_x, max_x, min_y, max_y, N):
    cands = np.ascontiguousarray(cands, CAND_DT)
    out = np.zeros(max(len(cands), 1), np.int32); n = C.c_int(0)
    rc = lib().ref_distribute(_p(cands), len(cands), min_x, max_x, min_y, max_y, N, _p(out), len(out), C.byref(n))
    if rc:
        raise RuntimeError(f"ref_distribute rc={rc}")
    return out[:n.value].copy()


def ic_angle(img, xs, ys):
    img = np.ascontiguousarray(img, np.uint8)
    xs = np.ascontiguousarray(xs, np.int32); ys = np.ascontiguousarray(ys, np.int32)
    out = np.zeros(len(xs), np.float32)
    lib().ref_ic_angle(_p(img), img.shape[1], img.shape[0], img.shape[1], _p(xs), _p(ys), len(xs), _p(out))
    return out


def orb_descriptor(blurred, xs, ys, angles):
    blurred = np.ascontiguousarray(blurred, np.uint8)
    xs = np.ascontiguousarray(xs, np.int32); ys = np.ascontiguousarray(ys, np.int32); angles = np.ascontiguousarray(angles, np.float32)
    out = np.zeros((len(xs), 32), np.uint8)
    lib().ref_orb_descriptor(_p(blurred), blurred.shape[1], blurred.shape[0], blurred.shape[1], _p(xs), _p(ys), _p(angles), len(xs), _p(out))
    return out


def tables(nfeatures=1000, scale_factor=1.2, nlevels=8):
    nf = np.zeros(nlevels, np.int32); um = np.zeros(16, np.int32)
    lib().ref_tables(nfeatures, C.c_float(scale_factor), nlevels, _p(nf), _p(um))
    return nf, um
