"""ctypes binding of the CPU oracle (oracle/liborb_oracle.so) — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module.  The product path (adaptive-rgbd-localization-mappig_b200/) never does.
Each wrapper cites the reference lines its C function restates (see the .cpp headers).
"""
import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_DIR = Path(__file__).resolve().parent
MAX_LEVELS = 16

KEYPOINT_DT = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                        ("octave", "<i4"), ("class_id", "<i4")])
DMATCH_DT = np.dtype([("queryIdx", "<i4"), ("trainIdx", "<i4"), ("imgIdx", "<i4"), ("distance", "<f4")])
CAND_DT = np.dtype([("x", "<i4"), ("y", "<i4"), ("score", "<i4")])
HYP_DT = np.dtype([("n_refined", "<i4"), ("rounds", "<i4"), ("refined_error", "<f8"), ("T", "<f4", (16,))])
assert KEYPOINT_DT.itemsize == 28 and DMATCH_DT.itemsize == 16 and HYP_DT.itemsize == 80


class ExtractCfg(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("nfeatures", C.c_int), ("nlevels", C.c_int),
                ("scale_factor", C.c_float), ("ini_th_fast", C.c_int), ("min_th_fast", C.c_int)]


class ExtractDebug(C.Structure):
    _fields_ = [("pyramid", C.c_void_p), ("blurred", C.c_void_p), ("cands", C.c_void_p), ("cand_cap", C.c_int),
                ("n_cands", C.c_void_p), ("n_kps", C.c_void_p), ("level_xy", C.c_void_p)]


class RansacCfg(C.Structure):
    _fields_ = [("iterations", C.c_int), ("min_inlier_th", C.c_uint), ("max_mahal", C.c_float),
                ("sample_size", C.c_uint), ("check_depth", C.c_int), ("depth_cov", C.c_double)]


class RansacOut(C.Structure):
    _fields_ = [("ok", C.c_int), ("rmse", C.c_float), ("T12", C.c_float * 16), ("n_inliers", C.c_int),
                ("n_good", C.c_int), ("real_iters", C.c_int), ("valid_iters", C.c_int), ("used_identity", C.c_int),
                ("depth_cov_used", C.c_double)]


def build(speed=False):
    """Compile the oracle with oracle/Makefile (g++ only; building the checker is not using it)."""
    import fcntl
    with open(_DIR / ".build.lock", "w") as lock:          # processes of one node share the tree
        fcntl.flock(lock, fcntl.LOCK_EX)
        subprocess.run(["make", "-C", str(_DIR), "-s", "all"], check=True)


_libs = {}


def lib(speed=False):
    name = "liborb_oracle_speed.so" if speed else "liborb_oracle.so"
    if name not in _libs:
        path = _DIR / name
        if not path.exists():
            build()
        L = C.CDLL(str(path))
        L.orc_fast_atan2.restype = C.c_float
        L.orc_fast_atan2.argtypes = [C.c_float, C.c_float]
        L.orc_mahalanobis2.restype = C.c_double
        L.orc_pattern.restype = C.POINTER(C.c_int8)
        _libs[name] = L
    return _libs[name]


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _chk(rc, what):
    if rc != 0:
        raise RuntimeError(f"oracle {what} failed rc={rc}")


def tables(nfeatures=1000, scale_factor=1.2, nlevels=8):
    sc = np.zeros(nlevels, np.float32); isc = np.zeros_like(sc); s2 = np.zeros_like(sc); is2 = np.zeros_like(sc)
    nf = np.zeros(nlevels, np.int32); um = np.zeros(16, np.int32)
    _chk(lib().orc_tables(nfeatures, C.c_float(scale_factor), nlevels, _p(sc), _p(isc), _p(s2), _p(is2), _p(nf), _p(um)),
         "tables")
    return dict(scale=sc, inv_scale=isc, sigma2=s2, inv_sigma2=is2, nfeat=nf, umax=um)


def level_sizes(w, h, scale_factor=1.2, nlevels=8):
    ws = np.zeros(nlevels, np.int32); hs = np.zeros(nlevels, np.int32)
    _chk(lib().orc_level_sizes(w, h, C.c_float(scale_factor), nlevels, _p(ws), _p(hs)), "level_sizes")
    return ws, hs


def resize_linear(src, dw, dh):
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.zeros((dh, dw), np.uint8)
    _chk(lib().orc_resize_linear(_p(src), src.shape[1], src.shape[0], src.shape[1], _p(dst), dw, dh, dw), "resize")
    return dst


def pyramid(img, scale_factor=1.2, nlevels=8):
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    ws, hs = level_sizes(w, h, scale_factor, nlevels)
    out = np.zeros(int((ws.astype(np.int64) * hs).sum()), np.uint8)
    _chk(lib().orc_pyramid(_p(img), w, h, w, C.c_float(scale_factor), nlevels, _p(out)), "pyramid")
    return split_levels(out, ws, hs)


def split_levels(buf, ws, hs):
    levels, off = [], 0
    for lw, lh in zip(ws, hs):
        levels.append(buf[off:off + int(lw) * int(lh)].reshape(int(lh), int(lw)))
        off += int(lw) * int(lh)
    return levels


def fast_roi(roi, th):
    roi = np.ascontiguousarray(roi, np.uint8)
    cap = roi.size // 2 + 16
    out = np.zeros(cap, CAND_DT); n = C.c_int(0)
    _chk(lib().orc_fast_roi(_p(roi), roi.shape[1], roi.shape[1], roi.shape[0], th, _p(out), cap, C.byref(n)), "fast_roi")
    return out[:n.value].copy()


def fast_strength_map(img):
    img = np.ascontiguousarray(img, np.uint8)
    out = np.zeros(img.shape, np.int16)
    _chk(lib().orc_fast_strength_map(_p(img), img.shape[1], img.shape[0], img.shape[1], _p(out)), "strength")
    return out


def fast_cells(img, ini_th=20, min_th=7):
    img = np.ascontiguousarray(img, np.uint8)
    cap = img.size // 4 + 16
    out = np.zeros(cap, CAND_DT); n = C.c_int(0)
    _chk(lib().orc_fast_cells(_p(img), img.shape[1], img.shape[0], img.shape[1], ini_th, min_th, _p(out), cap,
                              C.byref(n)), "fast_cells")
    return out[:n.value].copy()


def distribute(cands, min_x, max_x, min_y, max_y, N):
    cands = np.ascontiguousarray(cands, CAND_DT)
    out = np.zeros(max(len(cands), 1), np.int32); n = C.c_int(0)
    _chk(lib().orc_distribute(_p(cands), len(cands), min_x, max_x, min_y, max_y, N, _p(out), len(out), C.byref(n)),
         "distribute")
    return out[:n.value].copy()


def fast_atan2(y, x):
    return float(lib().orc_fast_atan2(C.c_float(y), C.c_float(x)))


def ic_angle(img, xs, ys):
    img = np.ascontiguousarray(img, np.uint8)
    xs = np.ascontiguousarray(xs, np.int32); ys = np.ascontiguousarray(ys, np.int32)
    out = np.zeros(len(xs), np.float32)
    _chk(lib().orc_ic_angle(_p(img), img.shape[1], img.shape[0], img.shape[1], _p(xs), _p(ys), len(xs), _p(out)), "ic_angle")
    return out


def gaussian_blur7(img):
    img = np.ascontiguousarray(img, np.uint8)
    out = np.zeros_like(img)
    _chk(lib().orc_gaussian_blur7(_p(img), img.shape[1], img.shape[0], img.shape[1], _p(out), img.shape[1]), "blur")
    return out


def rbrief(blurred, xs, ys, angles):
    blurred = np.ascontiguousarray(blurred, np.uint8)
    xs = np.ascontiguousarray(xs, np.int32); ys = np.ascontiguousarray(ys, np.int32)
    angles = np.ascontiguousarray(angles, np.float32)
    out = np.zeros((len(xs), 32), np.uint8)
    _chk(lib().orc_rbrief(_p(blurred), blurred.shape[1], blurred.shape[0], blurred.shape[1], _p(xs), _p(ys), _p(angles),
                          len(xs), _p(out)), "rbrief")
    return out


def pattern():
    return np.ctypeslib.as_array(lib().orc_pattern(), shape=(1024,)).copy()


def extract(img, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7, debug=False, speed=False):
    """ORBextractor::operator() restated (orbextractor.cpp:756-815).  Returns (kps, desc[, dbg])."""
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    cfg = ExtractCfg(w, h, nfeatures, nlevels, scale_factor, ini_th, min_th)
    cap = nfeatures + 8 * nlevels + 64
    kps = np.zeros(cap, KEYPOINT_DT); desc = np.zeros((cap, 32), np.uint8); n = C.c_int(0)
    dbg = None; d = {}
    if debug:
        ws, hs = level_sizes(w, h, scale_factor, nlevels)
        tot = int((ws.astype(np.int64) * hs).sum())
        d = dict(ws=ws, hs=hs, pyramid=np.zeros(tot, np.uint8), blurred=np.zeros(tot, np.uint8),
                 cands=np.zeros(tot // 4 + 64, CAND_DT), n_cands=np.zeros(nlevels, np.int32),
                 n_kps=np.zeros(nlevels, np.int32), level_xy=np.zeros((cap, 2), np.int32))
        dbg = ExtractDebug(_p(d["pyramid"]), _p(d["blurred"]), _p(d["cands"]), len(d["cands"]), _p(d["n_cands"]),
                           _p(d["n_kps"]), _p(d["level_xy"]))
    rc = lib(speed).orc_extract(C.byref(cfg), _p(img), w, _p(kps), _p(desc), cap, C.byref(n),
                                C.byref(dbg) if dbg else None)
    _chk(rc, "extract")
    k = n.value
    if debug:
        d["cands"] = d["cands"][:int(d["n_cands"].sum())].copy()
        d["level_xy"] = d["level_xy"][:k].copy()
        return kps[:k].copy(), desc[:k].copy(), d
    return kps[:k].copy(), desc[:k].copy()


FR1 = dict(fx=517.3, fy=516.5, cx=318.6, cy=255.3, mbf=40.0, depth_factor=np.float32(1.0) / np.float32(5000.0))


def unproject(kps, depth, depth_factor=FR1["depth_factor"], fx=FR1["fx"], fy=FR1["fy"], cx=FR1["cx"], cy=FR1["cy"],
              mbf=FR1["mbf"], dist=None):
    """Frame::ExtractFeatures depth gather + unprojection (frame.cpp:148-164).  dist = (k1, k2, p1, p2, k3) with k1 != 0 runs
    Frame::UndistortKeyPoints first (frame.cpp:286-313) and unprojects mvKeysUn; the depth lookup stays at the distorted keypoint."""
    kps = np.ascontiguousarray(kps, KEYPOINT_DT)
    if dist is not None and float(dist[0]) != 0.0:
        xy = np.stack([kps["x"], kps["y"]], 1).astype(np.float32) if len(kps) else np.zeros((0, 2), np.float32)
        xy_un = undistort_points(xy, fx, fy, cx, cy, dist)
        h, w = depth.shape
        xyz = np.zeros((len(kps), 3), np.float32); ur = np.zeros(len(kps), np.float32)
        if depth.dtype == np.uint16:
            d = np.ascontiguousarray(depth); args = (_p(d), None)
        else:
            d = np.ascontiguousarray(depth, np.float32); args = (None, _p(d))
        _chk(lib().orc_unproject_un(_p(kps), _p(xy_un), len(kps), args[0], args[1], w, h, w, C.c_float(depth_factor), C.c_float(fx),
                                    C.c_float(fy), C.c_float(cx), C.c_float(cy), C.c_float(mbf), _p(xyz), _p(ur)), "unproject_un")
        return xyz, ur
    h, w = depth.shape
    xyz = np.zeros((len(kps), 3), np.float32); ur = np.zeros(len(kps), np.float32)
    if depth.dtype == np.uint16:
        d = np.ascontiguousarray(depth); args = (_p(d), None)
    else:
        d = np.ascontiguousarray(depth, np.float32); args = (None, _p(d))
    _chk(lib().orc_unproject(_p(kps), len(kps), args[0], args[1], w, h, w, C.c_float(depth_factor), C.c_float(fx),
                             C.c_float(fy), C.c_float(cx), C.c_float(cy), C.c_float(mbf), _p(xyz), _p(ur)), "unproject")
    return xyz, ur


def distinctive_descriptors(desc, offsets):
    """Landmark::ComputeDistinctiveDescriptors (landmark.cpp:219-273) per landmark: (best row index, its median distance)."""
    desc = np.ascontiguousarray(desc, np.uint8); offsets = np.ascontiguousarray(offsets, np.int32)
    n = len(offsets) - 1
    best = np.zeros(n, np.int32); med = np.zeros(n, np.int32)
    _chk(lib().orc_distinctive_descriptors(_p(desc), _p(offsets), n, _p(best), _p(med)), "distinctive_descriptors")
    return best, med


def bgr2gray(bgr):
    """Frame::Frame's cv::cvtColor(imColor, mImGray, CV_BGR2GRAY) (Core/frame.cpp:23) for 8-bit images: OpenCV's fixed-point
    path with 15 fractional bits (pinned against cv2 4.13.0 in tests/test_ingest.py)."""
    b, g, r = (bgr[..., i].astype(np.uint32) for i in range(3))
    return ((b * 3735 + g * 19235 + r * 9798 + 16384) >> 15).astype(np.uint8)


class AdaptiveCfg(C.Structure):
    _fields_ = [("min_features", C.c_int), ("max_features", C.c_int), ("max_iters", C.c_int), ("max_per_cell", C.c_int),
                ("grid", C.c_int), ("edge", C.c_int), ("init_th", C.c_double), ("min_th", C.c_double), ("max_th", C.c_double),
                ("inc", C.c_double), ("dec", C.c_double)]


def adaptive_default(**kw):
    cfg = AdaptiveCfg()
    lib().orc_adaptive_default(C.byref(cfg))
    for k, v in kw.items():
        setattr(cfg, k, v)
    return cfg


def adaptive_detect(img, thresh, retain_best=0, cfg=None):
    """VideoGridAdaptedFeatureDetector::detect (+ retainBest) on one frame; `thresh` [grid*grid] float64 is updated in place.
    Returns (kps, cell_found, cell_thresh)."""
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    cfg = cfg or adaptive_default()
    g2 = cfg.grid * cfg.grid
    assert thresh.dtype == np.float64 and len(thresh) == g2
    cap = cfg.max_per_cell * g2 + 16
    out = np.zeros(cap, KEYPOINT_DT); n = C.c_int(0)
    found = np.zeros(g2, np.int32); used = np.zeros(g2, np.int32)
    _chk(lib().orc_adaptive_detect(C.byref(cfg), _p(img), w, h, w, _p(thresh), int(retain_best), _p(out), cap, C.byref(n), _p(found),
                                   _p(used)), "adaptive_detect")
    return out[:n.value].copy(), found, used


def extract_adapted(img, thresh, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7, cfg=None):
    """BASELINE config 4, 8-level variant: ORB extraction with per-region adapted iniThFAST; `thresh` [grid*grid] float64 is updated in
    place.  cfg.min_features / max_features = per-region band.  Returns (kps, desc, region_th, region_found)."""
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    ecfg = ExtractCfg(w, h, nfeatures, nlevels, scale_factor, ini_th, min_th)
    cfg = cfg or adaptive_default()
    g2 = cfg.grid * cfg.grid
    assert thresh.dtype == np.float64 and len(thresh) == g2
    cap = 2 * nfeatures + 64 * nlevels
    kps = np.zeros(cap, KEYPOINT_DT); desc = np.zeros((cap, 32), np.uint8); n = C.c_int(0)
    used = np.zeros(g2, np.int32); found = np.zeros(g2, np.int32)
    _chk(lib().orc_extract_adapted(C.byref(ecfg), C.byref(cfg), _p(img), w, _p(thresh), _p(kps), _p(desc), cap, C.byref(n), _p(used), _p(found)),
         "extract_adapted")
    return kps[:n.value].copy(), desc[:n.value].copy(), used, found


def projection_match(kp_x, kp_y, kp_octave, desc, lm_desc, proj_x, proj_y, lm_flags, feat_taken=None, radius=8.0, nn_ratio=0.8, th_high=100.0):
    """Matcher::ProjectionMatch (Features/matcher.cpp:90-143): (best feature per landmark or -1, number of matches)."""
    kp_x = np.ascontiguousarray(kp_x, np.float32); kp_y = np.ascontiguousarray(kp_y, np.float32)
    kp_octave = np.ascontiguousarray(kp_octave, np.int32); desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
    lm_desc = np.ascontiguousarray(lm_desc, np.uint8).reshape(-1, 32)
    proj_x = np.ascontiguousarray(proj_x, np.float32); proj_y = np.ascontiguousarray(proj_y, np.float32)
    lm_flags = np.ascontiguousarray(lm_flags, np.uint8)
    taken = None if feat_taken is None else np.ascontiguousarray(feat_taken, np.uint8)
    n, L = len(kp_x), len(lm_flags)
    best = np.full(max(L, 1), -1, np.int32); nm = C.c_int(0)
    _chk(lib().orc_projection_match(_p(kp_x) if n else None, _p(kp_y) if n else None, _p(kp_octave) if n else None, _p(desc) if n else None, n,
                                    _p(lm_desc) if L else None, _p(proj_x) if L else None, _p(proj_y) if L else None, _p(lm_flags) if L else None, L,
                                    _p(taken) if taken is not None else None, C.c_float(radius), C.c_float(nn_ratio), C.c_double(th_high), _p(best),
                                    C.byref(nm)), "projection_match")
    return best[:L], nm.value


def fuse_search(Rcw, tcw, camera, kp_x, kp_y, u_right, desc, lm_pos, lm_desc, lm_valid, radius=3.0, th_low=50.0):
    """Matcher::Fuse, projection + windowed search (Features/matcher.cpp:212-296)."""
    Rcw = np.ascontiguousarray(Rcw, np.float32).reshape(9); tcw = np.ascontiguousarray(tcw, np.float32).reshape(3)
    fx, fy, cx, cy, mbf, x0, x1, y0, y1 = [float(v) for v in np.asarray(camera, np.float32)]
    kp_x = np.ascontiguousarray(kp_x, np.float32); kp_y = np.ascontiguousarray(kp_y, np.float32); u_right = np.ascontiguousarray(u_right, np.float32)
    desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
    lm_pos = np.ascontiguousarray(lm_pos, np.float32).reshape(-1, 3); lm_desc = np.ascontiguousarray(lm_desc, np.uint8).reshape(-1, 32)
    lm_valid = np.ascontiguousarray(lm_valid, np.uint8)
    n, L = len(kp_x), len(lm_valid)
    best = np.full(max(L, 1), -1, np.int32); dist = np.full(max(L, 1), -1, np.int32)
    opt = lambda a: _p(a) if len(a) else None
    _chk(lib().orc_fuse_search(_p(Rcw), _p(tcw), C.c_float(fx), C.c_float(fy), C.c_float(cx), C.c_float(cy), C.c_float(mbf), C.c_float(x0), C.c_float(x1),
                               C.c_float(y0), C.c_float(y1), opt(kp_x), opt(kp_y), opt(u_right), opt(desc), n, opt(lm_pos), opt(lm_desc), opt(lm_valid), L,
                               C.c_float(radius), C.c_double(th_low), _p(best), _p(dist)), "fuse_search")
    return best[:L], dist[:L]


def bow_match(words1, off1, idx1, desc1, words2, off2, idx2, desc2, nn_ratio=0.6, th_low=50.0):
    """Matcher::BoWMatch (Features/matcher.cpp:145-209) on flattened feature vectors."""
    w1 = np.ascontiguousarray(words1, np.int32); o1 = np.ascontiguousarray(off1, np.int32); i1 = np.ascontiguousarray(idx1, np.int32)
    w2 = np.ascontiguousarray(words2, np.int32); o2 = np.ascontiguousarray(off2, np.int32); i2 = np.ascontiguousarray(idx2, np.int32)
    d1 = np.ascontiguousarray(desc1, np.uint8).reshape(-1, 32); d2 = np.ascontiguousarray(desc2, np.uint8).reshape(-1, 32)
    out = np.zeros(max(len(i1), 1), DMATCH_DT); n = C.c_int(0)
    opt = lambda a: _p(a) if len(a) else None
    _chk(lib().orc_bow_match(opt(w1), _p(o1), opt(i1), len(w1), opt(d1), opt(w2), _p(o2), opt(i2), len(w2), opt(d2), C.c_float(nn_ratio), C.c_double(th_low),
                             _p(out), len(out), C.byref(n)), "bow_match")
    return out[:n.value].copy()


def compose_trajectory(T12, pose0=None):
    """pose[k+1] = T12[k] * pose[k] (Odometry/odometry.cpp:82-84) for a whole sequence: [npairs + 1, 4, 4]."""
    T12 = np.ascontiguousarray(T12, np.float32).reshape(-1, 16)
    pose0 = np.eye(4, dtype=np.float32) if pose0 is None else np.ascontiguousarray(pose0, np.float32)
    out = np.zeros((len(T12) + 1, 16), np.float32)
    _chk(lib().orc_compose_trajectory(_p(T12) if len(T12) else None, len(T12), _p(pose0.reshape(16)), _p(out)), "compose_trajectory")
    return out.reshape(-1, 4, 4)


def undistort_points(xy, fx, fy, cx, cy, dist):
    """cv::undistortPoints(pts, pts, K, dist, Mat(), K) (Core/frame.cpp:302), dist = (k1, k2, p1, p2, k3)."""
    xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2); dist = np.ascontiguousarray(dist, np.float32).reshape(5)
    out = np.zeros_like(xy)
    _chk(lib().orc_undistort_points(_p(xy) if len(xy) else None, len(xy), C.c_float(fx), C.c_float(fy), C.c_float(cx), C.c_float(cy), _p(dist),
                                    _p(out) if len(xy) else None), "undistort_points")
    return out


def knn2(q, t, speed=False):
    q = np.ascontiguousarray(q, np.uint8); t = np.ascontiguousarray(t, np.uint8)
    nq, nt = len(q), len(t)
    o = [np.zeros(nq, np.int32) for _ in range(4)]
    _chk(lib(speed).orc_knn2(_p(q), nq, _p(t), nt, _p(o[0]), _p(o[1]), _p(o[2]), _p(o[3])), "knn2")
    return tuple(o)


def knn_match(q, t, ratio, cross_check=False, speed=False):
    q = np.ascontiguousarray(q, np.uint8); t = np.ascontiguousarray(t, np.uint8)
    out = np.zeros(max(len(q), 1), DMATCH_DT); n = C.c_int(0)
    _chk(lib(speed).orc_knn_match(_p(q), len(q), _p(t), len(t), C.c_float(ratio), int(cross_check), _p(out), len(out),
                                  C.byref(n)), "knn_match")
    return out[:n.value].copy()


def sample_table_libc(seed, M, iterations=200, sample_size=4):
    t = np.zeros((iterations, sample_size), np.int32)
    _chk(lib().orc_sample_table_libc(C.c_uint(seed), M, iterations, sample_size, _p(t)), "sample_table")
    return t


def libc_rand_sequence(seed, n):
    o = np.zeros(n, np.int32)
    _chk(lib().orc_libc_rand_sequence(C.c_uint(seed), n, _p(o)), "rand")
    return o


def std_sort_dmatch(m):
    m = np.ascontiguousarray(m, DMATCH_DT).copy()
    _chk(lib().orc_std_sort_dmatch(_p(m), len(m)), "sort")
    return m


def svd3(A):
    A = np.ascontiguousarray(A, np.float32)
    U = np.zeros((3, 3), np.float32); S = np.zeros(3, np.float32); V = np.zeros((3, 3), np.float32)
    _chk(lib().orc_svd3(_p(A), _p(U), _p(S), _p(V)), "svd3")
    return U, S, V


def weighted_transform(src, dst):
    src = np.ascontiguousarray(src, np.float32); dst = np.ascontiguousarray(dst, np.float32)
    T = np.zeros(16, np.float32)
    _chk(lib().orc_weighted_transform(_p(src), _p(dst), len(src), _p(T)), "weighted_transform")
    return T.reshape(4, 4)


def mahalanobis2(p1, p2, T, depth_cov):
    p1 = np.ascontiguousarray(p1, np.float32); p2 = np.ascontiguousarray(p2, np.float32)
    T = np.ascontiguousarray(T, np.float32)
    return float(lib().orc_mahalanobis2(_p(p1), _p(p2), _p(T), C.c_double(depth_cov)))


def set_sum_order(tree):
    """Sensitivity probe: evaluate the 3- / 4-term inner sums of the PCL / Eigen restatements as Eigen's balanced tree (True) instead of
    left to right (False, the default and what the CUDA path implements).  Returns the previous setting."""
    return bool(lib().orc_set_sum_order(int(bool(tree))))


def kabsch(A, B):
    A = np.ascontiguousarray(A, np.float32); B = np.ascontiguousarray(B, np.float32)
    T = np.zeros(16, np.float32)
    _chk(lib().orc_kabsch(_p(A), _p(B), len(A), _p(T)), "kabsch")
    return T.reshape(4, 4)


def ransac_iterate(src_xyz, dst_xyz, m12, iterations=200, min_inlier_th=20, max_mahal=3.0, sample_size=4,
                   check_depth=True, depth_cov=-1.0, sort_mode=0, sample_table=None, seed=42, speed=False):
    """Ransac::Iterate(F1, F2, m12) (ransac.cpp:155-267).  Returns a dict."""
    src = np.ascontiguousarray(src_xyz, np.float32); dst = np.ascontiguousarray(dst_xyz, np.float32)
    m12 = np.ascontiguousarray(m12, DMATCH_DT)
    cfg = RansacCfg(iterations, min_inlier_th, max_mahal, sample_size, int(check_depth), depth_cov)
    out = RansacOut()
    inl = np.zeros(max(len(m12), 1), DMATCH_DT)
    hyp = np.zeros(iterations, HYP_DT)
    good = np.zeros(max(len(m12), 1), DMATCH_DT)
    tab_out = np.full((iterations, sample_size), -1, np.int32)
    tab = None if sample_table is None else np.ascontiguousarray(sample_table, np.int32)
    rc = lib(speed).orc_ransac_iterate(C.byref(cfg), _p(src), len(src), _p(dst), len(dst), _p(m12), len(m12), sort_mode,
                                       _p(tab), C.c_uint(seed), _p(inl), len(inl), C.byref(out), _p(hyp), _p(good),
                                       _p(tab_out))
    _chk(rc, "ransac")
    return dict(ok=bool(out.ok), rmse=float(out.rmse), T12=np.array(out.T12, np.float32).reshape(4, 4),
                inliers=inl[:out.n_inliers].copy(), n_good=out.n_good, real_iters=out.real_iters,
                valid_iters=out.valid_iters, used_identity=bool(out.used_identity),
                depth_cov=float(out.depth_cov_used), hyp=hyp, good_sorted=good[:out.n_good].copy(),
                sample_table=tab_out)


def ransac_clouds(src_xyz, dst_xyz, m12, min_inlier_th=20, check_depth=True):
    """Ransac::mpSourceCloud / mpTargetCloud as Iterate leaves them (ransac.cpp:163-189): cleared; nothing more when m12 has fewer than
    min_inlier_th entries; otherwise one pcl::PointXYZ (x, y, z, 1) per match whose two depths are valid, in m12 order (before the sort).
    Plain numpy restatement (an index gather: nothing to compile)."""
    src = np.asarray(src_xyz, np.float32).reshape(-1, 3); dst = np.asarray(dst_xyz, np.float32).reshape(-1, 3)
    a = np.zeros((0, 4), np.float32)
    if len(m12) < min_inlier_th:
        return a, a.copy()
    s = src[m12["queryIdx"]]; t = dst[m12["trainIdx"]]
    keep = np.ones(len(m12), bool)
    if check_depth:
        keep = ~(np.isnan(s[:, 2]) | np.isnan(t[:, 2]) | (s[:, 2] <= 0) | (t[:, 2] <= 0))
    one = np.ones((int(keep.sum()), 1), np.float32)
    return np.hstack([s[keep], one]), np.hstack([t[keep], one])


def knn_match_keyframe(kf_desc, f2_desc, ratio, kf_landmarks, is_bad, f2_landmarks):
    """Matcher::KnnMatch(KeyFrame*, Frame&, vMatches12) (matcher.cpp:23-53): ratio survivors in query order, kept when the keyframe holds a
    landmark at queryIdx (kf_landmarks[i] != 0) that is not bad and the frame's feature at trainIdx is still free; an accepted match
    hands the landmark to the frame (f2_landmarks is updated in place) — so a later match onto the same feature is skipped."""
    out = []
    for m in knn_match(kf_desc, f2_desc, ratio, False):
        lm = kf_landmarks[m["queryIdx"]]
        if lm == 0 or is_bad(lm) or f2_landmarks[m["trainIdx"]] != 0:
            continue
        f2_landmarks[m["trainIdx"]] = lm
        out.append(m)
    return np.array(out, DMATCH_DT) if out else np.zeros(0, DMATCH_DT)
