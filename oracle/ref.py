"""ctypes binding of oracle/_ref/*.so — TEST INFRASTRUCTURE ONLY.

liborb_ref.so is the reference's own Features/orbextractor.cpp (+ matcher.cpp, extractor.cpp and the adaptive detector sources),
libodometry_ref.so its Odometry/ransac.cpp + Odometry/kabsch.cpp, libframe_ref.so its Core/frame.cpp + keyframe.cpp + landmark.cpp + map.cpp,
compiled verbatim from /root/reference (oracle/Makefile, target _ref) against the OpenCV / Eigen / PCL stand-ins under oracle/ref_shim/.  Only tests/, bench.py's CPU legs and __graft_entry__.build() touch it.
It exists where the reference checkout exists (the authoring container); the prebuilt file travels to the GPU box with the
snapshot (oracle/_ref/ is git-ignored, not gpurun-ignored).  available() says whether it can be used.
"""
import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

from .oracle import CAND_DT, DMATCH_DT, KEYPOINT_DT, RansacCfg, RansacOut, level_sizes, _p

_DIR = Path(__file__).resolve().parent
SO = _DIR / "_ref" / "liborb_ref.so"
SO_ODOMETRY = _DIR / "_ref" / "libodometry_ref.so"
SO_FRAME = _DIR / "_ref" / "libframe_ref.so"
UTILS_DEMO = _DIR / "_ref" / "ref_utils_demo"      # Utils/utils.cpp verbatim: the reference's own LoadImages as a small program
REFERENCE = Path(os.environ.get("ORB_REFERENCE_DIR", "/root/reference"))
_lib = None


def build():
    """Compile from the reference checkout when it is present; otherwise keep whatever prebuilt file is there."""
    if (REFERENCE / "Features" / "orbextractor.cpp").exists():
        subprocess.run(["make", "-C", str(_DIR), "-s", "-j4", "_ref", f"REF={REFERENCE}"], check=True)
    return SO.exists() and SO_ODOMETRY.exists() and SO_FRAME.exists()


def available():
    return (SO.exists() and SO_ODOMETRY.exists() and SO_FRAME.exists()) or build()


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


# ---- Odometry/ransac.cpp + Odometry/kabsch.cpp (libodometry_ref.so) --------------------------------------------------------------
_odo = None


def odometry_lib():
    global _odo
    if _odo is None:
        if not available():
            raise RuntimeError("oracle/_ref/libodometry_ref.so is not built and the reference checkout is absent")
        _odo = C.CDLL(str(SO_ODOMETRY))
        _odo.ref_depth_covariance.restype = C.c_double
        _odo.ref_depth_covariance.argtypes = [C.c_double]
        _odo.ref_inliers_and_error.restype = C.c_double
    return _odo


def depth_covariance(depth):
    """Ransac::DepthCovariance(depth) (ransac.cpp:416-421, quirk Q7): the FIRST call in the process — this one, or the first scored
    point of an earlier Iterate — fixes what every later call returns."""
    return float(odometry_lib().ref_depth_covariance(float(depth)))


def ransac_iterate(src_xyz, dst_xyz, m12, iterations=200, min_inlier_th=20, max_mahal=3.0, sample_size=4, check_depth=True, seed=42,
                   member_form=False):
    """The reference's Ransac::Iterate(F1, F2, m12) (ransac.cpp:155-267) or, member_form, Ransac(KF1, KF2, m12) + Iterate()
    (ransac.cpp:26-36, 44-153), after srand(seed)."""
    src = np.ascontiguousarray(src_xyz, np.float32); dst = np.ascontiguousarray(dst_xyz, np.float32)
    m12 = np.ascontiguousarray(m12, DMATCH_DT)
    cfg = RansacCfg(iterations, min_inlier_th, max_mahal, sample_size, int(check_depth), -1.0)
    out = RansacOut()
    inl = np.zeros(max(len(m12), 1), DMATCH_DT)
    cs = np.zeros((max(len(m12), 1), 3), np.float32); ct = np.zeros_like(cs); nc = C.c_int(0)
    rc = odometry_lib().ref_ransac_iterate(C.byref(cfg), _p(src), len(src), _p(dst), len(dst), _p(m12), len(m12), C.c_uint(seed),
                                           int(member_form), _p(inl), len(inl), C.byref(out), _p(cs), _p(ct), C.byref(nc))
    if rc:
        raise RuntimeError(f"ref_ransac_iterate rc={rc}")
    return dict(ok=bool(out.ok), rmse=float(out.rmse), T12=np.array(out.T12, np.float32).reshape(4, 4), inliers=inl[:out.n_inliers].copy(),
                n_good=out.n_good, cloud_src=cs[:nc.value].copy(), cloud_tgt=ct[:nc.value].copy())


def sample_table(seed, M, iterations=200, sample_size=4):
    """Ransac::SampleMatches (ransac.cpp:269-292) called `iterations` times after srand(seed) on M matches."""
    tab = np.full((iterations, sample_size), -1, np.int32)
    odometry_lib().ref_sample_table(C.c_uint(seed), int(M), int(iterations), int(sample_size), _p(tab))
    return tab


def inliers_and_error(src_xyz, dst_xyz, m12, T, max_mahal=3.0):
    """Ransac::ComputeInliersAndError (ransac.cpp:313-348): (error, positions of the inliers in m12)."""
    src = np.ascontiguousarray(src_xyz, np.float32); dst = np.ascontiguousarray(dst_xyz, np.float32)
    m12 = np.ascontiguousarray(m12, DMATCH_DT); T = np.ascontiguousarray(T, np.float32)
    pos = np.zeros(max(len(m12), 1), np.int32); n = C.c_int(0)
    e = odometry_lib().ref_inliers_and_error(_p(src), len(src), _p(dst), len(dst), _p(m12), len(m12), _p(T), C.c_float(max_mahal), _p(pos), C.byref(n))
    return float(e), pos[:n.value].copy()


def kabsch(A, B):
    """Kabsch::Compute(setA, setB) (kabsch.cpp:14-57)."""
    A = np.ascontiguousarray(A, np.float32).reshape(-1, 3); B = np.ascontiguousarray(B, np.float32).reshape(-1, 3)
    T = np.zeros((4, 4), np.float32)
    odometry_lib().ref_kabsch(_p(A), _p(B), len(A), _p(T))
    return T


# ---- adaptive detector chain (row a-17): Features/extractor.cpp + video*adaptedfeaturedetector.cpp + detectoradjuster.cpp (liborb_ref.so) ----
class AdaptiveExtractor:
    """The reference's Extractor(FAST, ORB, ADAPTIVE); extract(img) = Extractor::Extract on the next frame of the video."""

    def __init__(self):
        L = lib()
        L.ref_adaptive_create.restype = C.c_void_p
        L.ref_adaptive_destroy.argtypes = [C.c_void_p]
        self._h = C.c_void_p(L.ref_adaptive_create())

    def extract(self, img, cap=4096):
        img = np.ascontiguousarray(img, np.uint8)
        h, w = img.shape
        out = np.zeros(cap, KEYPOINT_DT); n = C.c_int(0); th = np.zeros(9, np.float64)
        rc = lib().ref_adaptive_extract(self._h, _p(img), w, h, w, _p(out), cap, C.byref(n), _p(th))
        if rc:
            raise RuntimeError(f"ref_adaptive_extract rc={rc}")
        return out[:n.value].copy(), th

    def close(self):
        if self._h:
            lib().ref_adaptive_destroy(self._h); self._h = None


# ---- Features/matcher.cpp (liborb_ref.so): both KnnMatch overloads, ProjectionMatch, BoWMatch, Fuse -------------------------------------
def _opt(a):
    return _p(a) if len(a) else None


def knn_match_frames(q, t, ratio):
    """Matcher(ratio).KnnMatch(Frame&, Frame&, matches) (matcher.cpp:55-88), landmark filters passing everything."""
    q = np.ascontiguousarray(q, np.uint8).reshape(-1, 32); t = np.ascontiguousarray(t, np.uint8).reshape(-1, 32)
    out = np.zeros(max(len(q), 1), DMATCH_DT); n = C.c_int(0)
    rc = lib().ref_knn_match_frames(_opt(q), len(q), _opt(t), len(t), C.c_float(ratio), _p(out), len(out), C.byref(n))
    if rc:
        raise RuntimeError(f"ref_knn_match_frames rc={rc}")
    return out[:n.value].copy()


def knn_match_keyframe(q, t, ratio, kf_landmarks, lm_bad, f2_landmarks):
    """Matcher(ratio).KnnMatch(KeyFrame*, Frame&, matches) (matcher.cpp:23-53); f2_landmarks (int32 ids, 0 = free) is updated in place."""
    q = np.ascontiguousarray(q, np.uint8).reshape(-1, 32); t = np.ascontiguousarray(t, np.uint8).reshape(-1, 32)
    kf = np.ascontiguousarray(kf_landmarks, np.int32); bad = np.ascontiguousarray(lm_bad, np.uint8)
    assert f2_landmarks.dtype == np.int32 and f2_landmarks.flags.c_contiguous
    out = np.zeros(max(len(q), 1), DMATCH_DT); n = C.c_int(0)
    rc = lib().ref_knn_match_keyframe(_opt(q), len(q), _opt(t), len(t), C.c_float(ratio), _p(kf), _p(bad), len(bad), _p(f2_landmarks), _p(out), len(out),
                                      C.byref(n))
    if rc:
        raise RuntimeError(f"ref_knn_match_keyframe rc={rc}")
    return out[:n.value].copy()


def projection_match(kp_x, kp_y, kp_octave, desc, lm_desc, proj_x, proj_y, lm_flags, feat_taken=None, radius=8.0, nn_ratio=0.8, th_high=100.0):
    """Matcher(nn_ratio).ProjectionMatch(frame, landmarks, radius) (matcher.cpp:90-143): (slot per landmark or -1, nmatches)."""
    kp_x = np.ascontiguousarray(kp_x, np.float32); kp_y = np.ascontiguousarray(kp_y, np.float32)
    kp_octave = np.ascontiguousarray(kp_octave, np.int32); desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
    lm_desc = np.ascontiguousarray(lm_desc, np.uint8).reshape(-1, 32)
    proj_x = np.ascontiguousarray(proj_x, np.float32); proj_y = np.ascontiguousarray(proj_y, np.float32)
    lm_flags = np.ascontiguousarray(lm_flags, np.uint8)
    taken = None if feat_taken is None else np.ascontiguousarray(feat_taken, np.uint8)
    n, L = len(kp_x), len(lm_flags)
    best = np.full(max(L, 1), -1, np.int32); nm = C.c_int(0)
    rc = lib().ref_projection_match(_opt(kp_x), _opt(kp_y), _opt(kp_octave), _opt(desc), n, _opt(lm_desc), _opt(proj_x), _opt(proj_y), _opt(lm_flags), L,
                                    _p(taken) if taken is not None else None, C.c_float(radius), C.c_float(nn_ratio), C.c_double(th_high), _p(best),
                                    C.byref(nm))
    if rc:
        raise RuntimeError(f"ref_projection_match rc={rc}")
    return best[:L], nm.value


def bow_match(words1, off1, idx1, desc1, words2, off2, idx2, desc2, nn_ratio=0.6, th_low=50.0):
    """Matcher(nn_ratio).BoWMatch(KF1, KF2, matches) (matcher.cpp:145-209)."""
    w1 = np.ascontiguousarray(words1, np.int32); o1 = np.ascontiguousarray(off1, np.int32); i1 = np.ascontiguousarray(idx1, np.int32)
    w2 = np.ascontiguousarray(words2, np.int32); o2 = np.ascontiguousarray(off2, np.int32); i2 = np.ascontiguousarray(idx2, np.int32)
    d1 = np.ascontiguousarray(desc1, np.uint8).reshape(-1, 32); d2 = np.ascontiguousarray(desc2, np.uint8).reshape(-1, 32)
    out = np.zeros(max(len(i1), 1), DMATCH_DT); n = C.c_int(0)
    rc = lib().ref_bow_match(_opt(w1), _p(o1), _opt(i1), len(w1), _opt(d1), len(d1), _opt(w2), _p(o2), _opt(i2), len(w2), _opt(d2), len(d2),
                             C.c_float(nn_ratio), C.c_double(th_low), _p(out), len(out), C.byref(n))
    if rc:
        raise RuntimeError(f"ref_bow_match rc={rc}")
    return out[:n.value].copy()


def fuse(Rcw, tcw, bounds, kp_x, kp_y, u_right, desc, lm_pos, lm_desc, lm_state, radius=3.0, th_low=50.0):
    """Matcher().Fuse(KF, landmarks, radius) (matcher.cpp:212-311) with the reference's compiled-in FR1 calibration; bounds = (min_x, max_x,
    min_y, max_y); lm_state 0 null / 1 valid / 2 bad / 3 already in the keyframe.  Returns (feature per landmark or -1, nFused)."""
    Rcw = np.ascontiguousarray(Rcw, np.float32).reshape(9); tcw = np.ascontiguousarray(tcw, np.float32).reshape(3)
    kp_x = np.ascontiguousarray(kp_x, np.float32); kp_y = np.ascontiguousarray(kp_y, np.float32); u_right = np.ascontiguousarray(u_right, np.float32)
    desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
    lm_pos = np.ascontiguousarray(lm_pos, np.float32).reshape(-1, 3); lm_desc = np.ascontiguousarray(lm_desc, np.uint8).reshape(-1, 32)
    lm_state = np.ascontiguousarray(lm_state, np.uint8)
    L = len(lm_state)
    best = np.full(max(L, 1), -1, np.int32); nf = C.c_int(0)
    x0, x1, y0, y1 = [C.c_float(float(v)) for v in bounds]
    rc = lib().ref_fuse(_p(Rcw), _p(tcw), x0, x1, y0, y1, _opt(kp_x), _opt(kp_y), _opt(u_right), _opt(desc), len(kp_x), _opt(lm_pos), _opt(lm_desc),
                        _opt(lm_state), L, C.c_float(radius), C.c_double(th_low), _p(best), C.byref(nf))
    if rc:
        raise RuntimeError(f"ref_fuse rc={rc}")
    return best[:L], nf.value


# ---- Core/frame.cpp + keyframe.cpp + landmark.cpp + map.cpp with the reference's real Core classes (libframe_ref.so) --------------------
_frame = None


def frame_lib():
    global _frame
    if _frame is None:
        if not available():
            raise RuntimeError("oracle/_ref/libframe_ref.so is not built and the reference checkout is absent")
        _frame = C.CDLL(str(SO_FRAME))
    return _frame


def frame_extract(bgr, depth_u16, cap=4096):
    """Frame(imColor, imDepth, t) + Frame::ExtractFeatures(Extractor(ORB_SLAM2, ORB_SLAM2, NORMAL)) (Core/frame.cpp:18-45, 135-170) with the
    calibration compiled into the reference (Utils/common.h: FR1 and its distortion)."""
    bgr = np.ascontiguousarray(bgr, np.uint8); depth = np.ascontiguousarray(depth_u16, np.uint16)
    h, w = depth.shape
    assert bgr.shape == (h, w, 3)
    kps = np.zeros(cap, KEYPOINT_DT); desc = np.zeros((cap, 32), np.uint8); un = np.zeros((cap, 2), np.float32)
    xyz = np.zeros((cap, 3), np.float32); ur = np.zeros(cap, np.float32); n = C.c_int(0)
    gray = np.zeros((h, w), np.uint8); bounds = np.zeros(4, np.float32)
    rc = frame_lib().ref_frame_extract(_p(bgr), _p(depth), w, h, _p(kps), _p(desc), _p(un), _p(xyz), _p(ur), cap, C.byref(n), _p(gray), _p(bounds))
    if rc:
        raise RuntimeError(f"ref_frame_extract rc={rc}")
    k = n.value
    return dict(kps=kps[:k].copy(), desc=desc[:k].copy(), xy_un=un[:k].copy(), xyz=xyz[:k].copy(), uright=ur[:k].copy(), gray=gray, bounds=bounds)


def distinctive_descriptors(desc, offsets, bad=None):
    """Landmark::ComputeDistinctiveDescriptors (Core/landmark.cpp:219-273) per landmark: (descriptor [n, 32], set flag [n])."""
    desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32); offsets = np.ascontiguousarray(offsets, np.int32)
    n = len(offsets) - 1
    b = None if bad is None else np.ascontiguousarray(bad, np.uint8)
    out = np.zeros((max(n, 1), 32), np.uint8); has = np.zeros(max(n, 1), np.uint8)
    frame_lib().ref_distinctive_descriptors(_p(desc) if len(desc) else None, _p(b) if b is not None else None, _p(offsets), n, _p(out), _p(has))
    return out[:n], has[:n].astype(bool)


def frame_depth_covariance(depth):
    """Ransac::DepthCovariance of libframe_ref.so's copy of ransac.cpp (quirk Q7: first call in the process latches)."""
    L = frame_lib()
    L.ref_frame_depth_covariance.restype = C.c_double; L.ref_frame_depth_covariance.argtypes = [C.c_double]
    return float(L.ref_frame_depth_covariance(float(depth)))


def odometry_compute(src_xyz, dst_xyz, m12, pose1, seed=42):
    """Odometry(RANSAC).Compute(pF1, pF2, m12) (Odometry/odometry.cpp:44, 78-90) on the reference's real Frame objects after srand(seed)."""
    src = np.ascontiguousarray(src_xyz, np.float32); dst = np.ascontiguousarray(dst_xyz, np.float32)
    m12 = np.ascontiguousarray(m12, DMATCH_DT); p1 = np.ascontiguousarray(pose1, np.float32).reshape(16)
    p2 = np.zeros(16, np.float32); outl = np.zeros(max(len(dst), 1), np.uint8); out = RansacOut(); inl = np.zeros(max(len(m12), 1), DMATCH_DT)
    rc = frame_lib().ref_odometry_compute(_p(src), len(src), _p(dst), len(dst), _p(m12), len(m12), C.c_uint(seed), _p(p1), _p(p2), _p(outl), C.byref(out),
                                          _p(inl), len(inl))
    if rc:
        raise RuntimeError(f"ref_odometry_compute rc={rc}")
    return dict(ok=bool(out.ok), rmse=float(out.rmse), T12=np.array(out.T12, np.float32).reshape(4, 4), inliers=inl[:out.n_inliers].copy(),
                pose2=p2.reshape(4, 4), outlier2=outl[:len(dst)].astype(bool))


def real_knn_match_frames(q, t, ratio, lm_obs1, outlier1, lm_obs2):
    """Matcher(ratio).KnnMatch(Frame&, Frame&, matches) (matcher.cpp:55-88) on the reference's real Frame / Landmark objects:
    (matches, slot2 [nt] = F1 feature whose landmark ended in F2's slot or -1, outlier2 [nt])."""
    q = np.ascontiguousarray(q, np.uint8).reshape(-1, 32); t = np.ascontiguousarray(t, np.uint8).reshape(-1, 32)
    o1 = np.ascontiguousarray(lm_obs1, np.int32); x1 = np.ascontiguousarray(outlier1, np.uint8); o2 = np.ascontiguousarray(lm_obs2, np.int32)
    out = np.zeros(max(len(q), 1), DMATCH_DT); n = C.c_int(0); slot2 = np.zeros(max(len(t), 1), np.int32); outl2 = np.zeros(max(len(t), 1), np.uint8)
    rc = frame_lib().ref_real_knn_match_frames(_p(q), len(q), _p(t), len(t), C.c_float(ratio), _p(o1), _p(x1), _p(o2), _p(out), len(out), C.byref(n),
                                               _p(slot2), _p(outl2))
    if rc:
        raise RuntimeError(f"ref_real_knn_match_frames rc={rc}")
    return out[:n.value].copy(), slot2[:len(t)], outl2[:len(t)].astype(bool)
