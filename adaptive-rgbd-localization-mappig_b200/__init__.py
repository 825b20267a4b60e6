"""Host-side Python mirror of the C ABI in include/orbfront.h (ctypes over liborbfront_b200.so).

The directory name is not a valid Python identifier; load it with
    importlib.util.spec_from_file_location("orbfront_b200", ".../adaptive-rgbd-localization-mappig_b200/__init__.py")
(tests/conftest.py, bench.py and __graft_entry__.py do exactly that).  Everything here is plumbing: numpy
arrays in, numpy arrays out, every call lands in a CUDA kernel of the shared library.  There is no CPU
fallback: if the library is missing it is built with nvcc, and if no CUDA device is present Context()
raises OrbfError(ORBF_ERR_CUDA).
"""
import ctypes as C
import importlib.util
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "liborbfront_b200.so"

KEYPOINT_DT = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                        ("octave", "<i4"), ("class_id", "<i4")])
DMATCH_DT = np.dtype([("queryIdx", "<i4"), ("trainIdx", "<i4"), ("imgIdx", "<i4"), ("distance", "<f4")])
CAND_DT = np.dtype([("x", "<i4"), ("y", "<i4"), ("score", "<i4")])
HYP_DT = np.dtype([("n_refined", "<i4"), ("rounds", "<i4"), ("refined_error", "<f8"), ("T", "<f4", (16,))])
RANSAC_RESULT_DT = np.dtype([("ok", "<i4"), ("rmse", "<f4"), ("T12", "<f4", (16,)), ("n_inliers", "<i4"), ("n_good", "<i4"),
                             ("real_iters", "<i4"), ("valid_iters", "<i4"), ("used_identity", "<i4"), ("_pad", "<i4"),
                             ("depth_cov_used", "<f8")])

STATUS = {0: "ORBF_OK", 1: "ORBF_ERR_ARG", 2: "ORBF_ERR_CAPACITY", 3: "ORBF_ERR_GEOMETRY", 4: "ORBF_ERR_CUDA",
          5: "ORBF_ERR_ALIGNMENT", 6: "ORBF_ERR_STATE"}


class Config(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("nfeatures", C.c_int32), ("nlevels", C.c_int32),
                ("scale_factor", C.c_float), ("ini_th_fast", C.c_int32), ("min_th_fast", C.c_int32),
                ("max_frames", C.c_int32), ("max_pairs", C.c_int32), ("device", C.c_int32),
                ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float), ("mbf", C.c_float),
                ("depth_factor", C.c_float), ("pipeline_chunk", C.c_int32), ("pipeline_streams", C.c_int32),
                ("depth_zero_copy", C.c_int32), ("pipeline_overlap", C.c_int32),
                ("k1", C.c_float), ("k2", C.c_float), ("p1", C.c_float), ("p2", C.c_float), ("k3", C.c_float)]


class RansacConfig(C.Structure):
    _fields_ = [("iterations", C.c_int32), ("min_inlier_th", C.c_uint32), ("max_mahal", C.c_float),
                ("sample_size", C.c_uint32), ("check_depth", C.c_int32), ("sort_mode", C.c_int32),
                ("depth_cov", C.c_double), ("seed", C.c_uint32)]


class AdaptiveConfig(C.Structure):
    _fields_ = [("min_features", C.c_int32), ("max_features", C.c_int32), ("max_iters", C.c_int32), ("max_per_cell", C.c_int32),
                ("grid", C.c_int32), ("edge", C.c_int32), ("init_th", C.c_double), ("min_th", C.c_double), ("max_th", C.c_double),
                ("inc", C.c_double), ("dec", C.c_double), ("retain_best", C.c_int32)]


class RansacResult(C.Structure):
    _fields_ = [("ok", C.c_int32), ("rmse", C.c_float), ("T12", C.c_float * 16), ("n_inliers", C.c_int32),
                ("n_good", C.c_int32), ("real_iters", C.c_int32), ("valid_iters", C.c_int32), ("used_identity", C.c_int32),
                ("depth_cov_used", C.c_double)]


assert C.sizeof(RansacResult) == RANSAC_RESULT_DT.itemsize, (C.sizeof(RansacResult), RANSAC_RESULT_DT.itemsize)


class OrbfError(RuntimeError):
    def __init__(self, status, what, detail=""):
        self.status = status
        super().__init__(f"{what}: {STATUS.get(status, status)} {detail}".strip())


def build(force=False):
    spec = importlib.util.spec_from_file_location("_orbf_build", _HERE / "build.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.build(force=force)


_lib = None


def lib():
    """Load (building first if needed) the shared library.  Raises if it cannot be produced: no fallback."""
    global _lib
    if _lib is None:
        import os
        alt = os.environ.get("ORBF_LIB")             # tuning runs: load another build of the same sources (tools/fast_variants.sh)
        if alt:
            L = C.CDLL(alt)
        else:
            build()
            L = C.CDLL(str(LIB_PATH))
        L.orbf_status_string.restype = C.c_char_p
        L.orbf_last_error.restype = C.c_char_p
        L.orbf_last_error.argtypes = [C.c_void_p]
        L.orbf_profile_stage_name.restype = C.c_char_p
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def default_config(**kw):
    cfg = Config()
    lib().orbf_default_config(C.byref(cfg))
    for k, v in kw.items():
        setattr(cfg, k, v)
    return cfg


def default_ransac_config(**kw):
    cfg = RansacConfig()
    lib().orbf_default_ransac_config(C.byref(cfg))
    for k, v in kw.items():
        setattr(cfg, k, v)
    return cfg


def descriptor_distance(a, b):
    """Matcher::DescriptorDistance (matcher.cpp:355-358) for NORM_HAMMING rows."""
    a = np.ascontiguousarray(a, np.uint8).ravel(); b = np.ascontiguousarray(b, np.uint8).ravel()
    d = C.c_int32(0)
    rc = lib().orbf_descriptor_distance(_p(a), _p(b), len(a), C.byref(d))
    if rc:
        raise OrbfError(rc, "descriptor_distance")
    return d.value


class Context:
    """One per GPU.  Mirrors ORBextractor + Matcher + Ransac + Kabsch behind the C ABI."""

    def __init__(self, **kw):
        self.cfg = default_config(**kw)
        self._h = C.c_void_p(None)
        rc = lib().orbf_create(C.byref(self.cfg), C.byref(self._h))
        if rc:
            detail = lib().orbf_last_error(self._h).decode() if self._h else ""
            if self._h:
                lib().orbf_destroy(self._h)
                self._h = C.c_void_p(None)
            raise OrbfError(rc, "orbf_create", detail)
        cap = C.c_int32(0)
        self._chk(lib().orbf_keypoint_capacity(self._h, C.byref(cap)), "keypoint_capacity")
        self.K = cap.value
        self.L = self.cfg.nlevels

    def close(self):
        if self._h:
            lib().orbf_destroy(self._h)
            self._h = C.c_void_p(None)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc, what):
        if rc:
            raise OrbfError(rc, what, lib().orbf_last_error(self._h).decode() if self._h else "")

    # ---- plumbing ----
    def set_stream(self, cuda_stream_ptr):
        self._chk(lib().orbf_set_stream(self._h, C.c_void_p(cuda_stream_ptr)), "set_stream")

    def synchronize(self):
        self._chk(lib().orbf_synchronize(self._h), "synchronize")

    def launch_count(self):
        n = C.c_int64(0)
        self._chk(lib().orbf_launch_count(self._h, C.byref(n)), "launch_count")
        return n.value

    def profile_enable(self, on=True):
        self._chk(lib().orbf_profile_enable(self._h, int(on)), "profile_enable")

    def profile_collect(self):
        self._chk(lib().orbf_profile_collect(self._h), "profile_collect")

    def profile_read(self):
        ms = np.zeros(10, np.float64); calls = np.zeros(10, np.int64)
        self._chk(lib().orbf_profile_read(self._h, _p(ms), _p(calls), 10), "profile_read")
        return {lib().orbf_profile_stage_name(i).decode(): (float(ms[i]), int(calls[i])) for i in range(10)}

    def tables(self):
        L = self.L
        f = [np.zeros(L, np.float32) for _ in range(4)]
        i = [np.zeros(L, np.int32) for _ in range(3)]
        self._chk(lib().orbf_get_tables(self._h, _p(f[0]), _p(f[1]), _p(f[2]), _p(f[3]), _p(i[0]), _p(i[1]), _p(i[2])), "get_tables")
        return dict(scale=f[0], inv_scale=f[1], sigma2=f[2], inv_sigma2=f[3], nfeat=i[0], level_w=i[1], level_h=i[2])

    # ---- extraction ----
    def extract(self, img):
        """ORBextractor::operator()(image, mask, keypoints, descriptors): host in, host out."""
        img = np.ascontiguousarray(img, np.uint8)
        h, w = img.shape if img.ndim == 2 else (0, 0)
        kps = np.zeros(self.K, KEYPOINT_DT); desc = np.zeros((self.K, 32), np.uint8); n = C.c_int32(0)
        self._chk(lib().orbf_extract(self._h, _p(img) if img.size else None, w, h, w, _p(kps), _p(desc), self.K, C.byref(n)), "extract")
        return kps[:n.value].copy(), desc[:n.value].copy()

    def extract_batch(self, frames, depths=None, slot0=0):
        """frames: [n, H, W] u8 host array (pinned or pageable); depths: [n, H, W] u16 or None.  Asynchronous."""
        assert frames.dtype == np.uint8 and frames.ndim == 3 and frames.flags.c_contiguous
        n, h, w = frames.shape
        if depths is not None:
            assert depths.dtype == np.uint16 and depths.shape == frames.shape and depths.flags.c_contiguous
        self._chk(lib().orbf_extract_batch(self._h, slot0, n, _p(frames), C.c_int64(w), C.c_int64(w * h), _p(depths),
                                           C.c_int64(w), C.c_int64(w * h)), "extract_batch")

    def extract_batch_bgr(self, bgr, depths=None, slot0=0):
        """bgr: [n, H, W, 3] u8 host frames (Frame::Frame's cvtColor runs on the device); depths as in extract_batch."""
        assert bgr.dtype == np.uint8 and bgr.ndim == 4 and bgr.shape[3] == 3 and bgr.flags.c_contiguous
        n, h, w, _ = bgr.shape
        if depths is not None:
            assert depths.dtype == np.uint16 and depths.shape == bgr.shape[:3] and depths.flags.c_contiguous
        self._chk(lib().orbf_extract_batch_bgr(self._h, slot0, n, _p(bgr), C.c_int64(3 * w), C.c_int64(3 * w * h), _p(depths), C.c_int64(w),
                                               C.c_int64(w * h)), "extract_batch_bgr")

    def download_gray(self, slot):
        out = np.zeros((self.cfg.height, self.cfg.width), np.uint8)
        self._chk(lib().orbf_download_gray(self._h, slot, _p(out), self.cfg.width), "download_gray")
        return out

    def extract_batch_device(self, d_gray_ptr, pitch, frame_stride, n, d_depth_ptr=0, depth_pitch=0, depth_frame_stride=0, slot0=0):
        self._chk(lib().orbf_extract_batch_device(self._h, slot0, n, C.c_void_p(d_gray_ptr), C.c_int64(pitch), C.c_int64(frame_stride),
                                                  C.c_void_p(d_depth_ptr) if d_depth_ptr else None, C.c_int64(depth_pitch),
                                                  C.c_int64(depth_frame_stride)), "extract_batch_device")

    def track_sequence(self, frames, depths, ratio, cross_check=False, ransac=True, slot0=0, **kw):
        """Tracking::Track over a host sequence: extract all frames, match + RANSAC consecutive pairs (pipelined, async)."""
        assert frames.dtype == np.uint8 and frames.ndim == 3 and frames.flags.c_contiguous
        n, h, w = frames.shape
        if depths is not None:
            assert depths.dtype == np.uint16 and depths.shape == frames.shape and depths.flags.c_contiguous
        cfg = default_ransac_config(**kw)
        self._chk(lib().orbf_track_sequence(self._h, slot0, n, _p(frames), C.c_int64(w), C.c_int64(w * h), _p(depths), C.c_int64(w),
                                            C.c_int64(w * h), C.c_float(ratio), int(cross_check), C.byref(cfg) if ransac else None),
                  "track_sequence")
        return n - 1

    def track_sequence_at(self, frames, depths, ratio, slot0, pair_slot0, cross_check=False, ransac=True, **kw):
        """track_sequence into frame slots [slot0, slot0 + n) and pair slots [pair_slot0, pair_slot0 + n - 1): with
        pipeline_overlap=1 two sequences can be double-buffered in one context."""
        n, h, w = frames.shape
        cfg = default_ransac_config(**kw)
        self._chk(lib().orbf_track_sequence_at(self._h, slot0, pair_slot0, n, _p(frames), C.c_int64(w), C.c_int64(w * h), _p(depths), C.c_int64(w),
                                               C.c_int64(w * h), C.c_float(ratio), int(cross_check), C.byref(cfg) if ransac else None),
                  "track_sequence_at")
        return n - 1

    def read_results_async(self, slot0, pair_slot0, n, frame_counts, match_counts, ransac, marker):
        """Asynchronous D2H of frame counts [n] i32, match counts [n - 1] i32 and RANSAC results [n - 1] RANSAC_RESULT_DT into caller
        arrays (page-locked for a truly asynchronous copy), then marker `marker`; wait_marker(marker) blocks until they have arrived."""
        self._chk(lib().orbf_read_results_async(self._h, slot0, pair_slot0, n, _p(frame_counts), _p(match_counts), _p(ransac), marker), "read_results_async")

    def read_features_async(self, slot0, pair_slot0, n, kps, desc, xyz, matches, marker):
        """Asynchronous D2H of kps [n, K] KEYPOINT_DT, desc [n, K, 32] u8, xyz [3, n, K] f32, matches [n - 1, K] DMATCH_DT (any may be None)."""
        self._chk(lib().orbf_read_features_async(self._h, slot0, pair_slot0, n, _p(kps), _p(desc), _p(xyz), _p(matches), marker), "read_features_async")

    def wait_marker(self, marker):
        self._chk(lib().orbf_wait_marker(self._h, marker), "wait_marker")

    def track_sequence_device(self, d_gray_ptr, pitch, frame_stride, n, d_depth_ptr, depth_pitch, depth_frame_stride, ratio,
                              cross_check=False, ransac=True, slot0=0, pair_slot0=0, **kw):
        """Device-resident sequence.  With pipeline_overlap=1 the call's RANSAC runs on a side stream: alternate disjoint (slot0, pair_slot0)
        halves between consecutive calls and they overlap; join() / synchronize() / any download orders the context stream behind it."""
        cfg = default_ransac_config(**kw)
        self._chk(lib().orbf_track_sequence_device_at(self._h, slot0, pair_slot0, n, C.c_void_p(d_gray_ptr), C.c_int64(pitch), C.c_int64(frame_stride),
                                                      C.c_void_p(d_depth_ptr) if d_depth_ptr else None, C.c_int64(depth_pitch),
                                                      C.c_int64(depth_frame_stride), C.c_float(ratio), int(cross_check),
                                                      C.byref(cfg) if ransac else None), "track_sequence_device")
        return n - 1

    def join(self):
        self._chk(lib().orbf_join(self._h), "join")

    def frame_counts(self, n, slot0=0):
        out = np.zeros(n, np.int32)
        self._chk(lib().orbf_frame_counts(self._h, slot0, n, _p(out)), "frame_counts")
        return out

    def download_frame(self, slot):
        kps = np.zeros(self.K, KEYPOINT_DT); desc = np.zeros((self.K, 32), np.uint8); xyz = np.zeros((self.K, 3), np.float32)
        n = C.c_int32(0)
        self._chk(lib().orbf_download_frame(self._h, slot, _p(kps), _p(desc), _p(xyz), self.K, C.byref(n)), "download_frame")
        return kps[:n.value].copy(), desc[:n.value].copy(), xyz[:n.value].copy()

    def download_keys_un(self, slot):
        """Frame::mvKeysUn (undistorted keypoint positions, [n, 2]) and mvuRight of a slot."""
        xy = np.zeros((self.K, 2), np.float32); ur = np.zeros(self.K, np.float32); n = C.c_int32(0)
        self._chk(lib().orbf_download_keys_un(self._h, slot, _p(xy), _p(ur), self.K, C.byref(n)), "download_keys_un")
        return xy[:n.value].copy(), ur[:n.value].copy()

    def pyramid_level(self, slot, level, blurred=False):
        t = self.tables()
        w, h = int(t["level_w"][level]), int(t["level_h"][level])
        out = np.zeros((h, w), np.uint8)
        self._chk(lib().orbf_pyramid_level(self._h, slot, level, int(blurred), _p(out), w), "pyramid_level")
        return out

    def level_candidates(self, slot, level):
        n = C.c_int32(0)
        rc = lib().orbf_level_candidates(self._h, slot, level, None, 0, C.byref(n))
        if rc not in (0, 2):
            self._chk(rc, "level_candidates")
        out = np.zeros(max(n.value, 1), CAND_DT)
        self._chk(lib().orbf_level_candidates(self._h, slot, level, _p(out), len(out), C.byref(n)), "level_candidates")
        return out[:n.value].copy()

    def level_keypoint_counts(self, slot):
        out = np.zeros(self.L, np.int32)
        self._chk(lib().orbf_level_keypoint_counts(self._h, slot, _p(out)), "level_keypoint_counts")
        return out

    # ---- matching ----
    def knn2(self, q, t):
        q = np.ascontiguousarray(q, np.uint8); t = np.ascontiguousarray(t, np.uint8)
        o = [np.full(len(q), -1, np.int32) for _ in range(4)]
        self._chk(lib().orbf_knn2(self._h, _p(q), len(q), _p(t), len(t), _p(o[0]), _p(o[1]), _p(o[2]), _p(o[3])), "knn2")
        return tuple(o)

    def knn_match(self, q, t, ratio, cross_check=False):
        q = np.ascontiguousarray(q, np.uint8); t = np.ascontiguousarray(t, np.uint8)
        out = np.zeros(max(len(q), 1), DMATCH_DT); n = C.c_int32(0)
        self._chk(lib().orbf_knn_match(self._h, _p(q), len(q), _p(t), len(t), C.c_float(ratio), int(cross_check), _p(out), len(out),
                                       C.byref(n)), "knn_match")
        return out[:n.value].copy()

    def distinctive_descriptors(self, desc, offsets):
        """Landmark::ComputeDistinctiveDescriptors for a batch of landmarks: (best row per landmark, its median distance)."""
        desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32); offsets = np.ascontiguousarray(offsets, np.int32)
        n = len(offsets) - 1
        best = np.zeros(max(n, 1), np.int32); med = np.zeros(max(n, 1), np.int32)
        self._chk(lib().orbf_distinctive_descriptors(self._h, _p(desc) if len(desc) else None, _p(offsets), n, _p(best), _p(med)),
                  "distinctive_descriptors")
        return best[:n], med[:n]

    def projection_match(self, lm_desc, proj_x, proj_y, lm_flags, slot=-1, kp_x=None, kp_y=None, kp_octave=None, desc=None, feat_taken=None,
                         radius=8.0, nn_ratio=0.8, th_high=100):
        """Matcher::ProjectionMatch (Features/matcher.cpp:90-143): (feature matched to each landmark or -1, number of matches).
        slot >= 0 uses that frame slot's keypoints / descriptors on the device; otherwise pass kp_x, kp_y, kp_octave, desc."""
        lm_desc = np.ascontiguousarray(lm_desc, np.uint8).reshape(-1, 32)
        proj_x = np.ascontiguousarray(proj_x, np.float32); proj_y = np.ascontiguousarray(proj_y, np.float32)
        lm_flags = np.ascontiguousarray(lm_flags, np.uint8)
        L = len(lm_flags)
        n = 0
        if slot < 0:
            kp_x = np.ascontiguousarray(kp_x, np.float32); kp_y = np.ascontiguousarray(kp_y, np.float32)
            kp_octave = np.ascontiguousarray(kp_octave, np.int32); desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
            n = len(kp_x)
        taken = None if feat_taken is None else np.ascontiguousarray(feat_taken, np.uint8)
        best = np.full(max(L, 1), -1, np.int32); nm = C.c_int32(0)
        opt = lambda a: _p(a) if a is not None and len(a) else None
        self._chk(lib().orbf_projection_match(self._h, int(slot), opt(kp_x) if slot < 0 else None, opt(kp_y) if slot < 0 else None,
                                              opt(kp_octave) if slot < 0 else None, opt(desc) if slot < 0 else None, n, opt(lm_desc), opt(proj_x), opt(proj_y),
                                              opt(lm_flags), L, opt(taken), C.c_float(radius), C.c_float(nn_ratio), int(th_high), _p(best), C.byref(nm)),
                  "projection_match")
        return best[:L], nm.value

    def compose_trajectory(self, npairs, pose0=None, with_flags=True):
        """Odometry::Compute's composition rule + inlier flags (Odometry/odometry.cpp:78-90) over the pairs last solved by ransac_pairs /
        track_sequence: (poses [npairs + 1, 4, 4], outlier flags [npairs + 1, K] or None)."""
        poses = np.zeros((npairs + 1, 16), np.float32)
        flags = np.ones((npairs + 1, self.K), np.uint8) if with_flags else None
        p0 = None if pose0 is None else np.ascontiguousarray(pose0, np.float32).reshape(16)
        self._chk(lib().orbf_compose_trajectory(self._h, int(npairs), _p(p0) if p0 is not None else None, _p(poses), _p(flags) if with_flags else None),
                  "compose_trajectory")
        return poses.reshape(-1, 4, 4), flags

    def undistort_points(self, xy, fx, fy, cx, cy, dist):
        """Frame::UndistortKeyPoints (Core/frame.cpp:286-313): cv::undistortPoints(pts, pts, K, dist, Mat(), K); dist = (k1, k2, p1, p2, k3)."""
        xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2); dist = np.ascontiguousarray(dist, np.float32).reshape(5)
        out = np.zeros_like(xy)
        self._chk(lib().orbf_undistort_points(self._h, _p(xy) if len(xy) else None, len(xy), C.c_float(fx), C.c_float(fy), C.c_float(cx), C.c_float(cy),
                                              _p(dist), _p(out) if len(xy) else None), "undistort_points")
        return out

    def unproject_keypoints(self, kps, depth):
        """Frame::ExtractFeatures' tail for external keypoints: (xyz [n, 3], u_right [n], xy_un [n, 2]); depth [H, W] u16 or None."""
        kps = np.ascontiguousarray(kps, KEYPOINT_DT); n = len(kps)
        xyz = np.zeros((n, 3), np.float32); ur = np.zeros(n, np.float32); un = np.zeros((n, 2), np.float32)
        if depth is not None:
            depth = np.ascontiguousarray(depth, np.uint16); h, w = depth.shape
        else:
            h, w = self.cfg.height, self.cfg.width
        self._chk(lib().orbf_unproject_keypoints(self._h, _p(kps), n, _p(depth), w, h, C.c_int64(w), _p(xyz), _p(ur), _p(un)), "unproject_keypoints")
        return xyz, ur, un

    def fuse_search(self, Rcw, tcw, camera, lm_pos, lm_desc, lm_valid, slot=-1, kp_x=None, kp_y=None, u_right=None, desc=None, radius=3.0, th_low=50):
        """Matcher::Fuse, projection + windowed search (Features/matcher.cpp:212-296): (best feature or -1, its distance or -1) per landmark.
        camera = (fx, fy, cx, cy, mbf, mnMinX, mnMaxX, mnMinY, mnMaxY)."""
        Rcw = np.ascontiguousarray(Rcw, np.float32).reshape(9); tcw = np.ascontiguousarray(tcw, np.float32).reshape(3)
        camera = np.ascontiguousarray(camera, np.float32).reshape(9)
        lm_pos = np.ascontiguousarray(lm_pos, np.float32).reshape(-1, 3); lm_desc = np.ascontiguousarray(lm_desc, np.uint8).reshape(-1, 32)
        lm_valid = np.ascontiguousarray(lm_valid, np.uint8)
        L = len(lm_valid); n = 0
        if slot < 0:
            kp_x = np.ascontiguousarray(kp_x, np.float32); kp_y = np.ascontiguousarray(kp_y, np.float32)
            u_right = np.ascontiguousarray(u_right, np.float32); desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
            n = len(kp_x)
        opt = lambda a: _p(a) if a is not None and len(a) else None
        best = np.full(max(L, 1), -1, np.int32); dist = np.full(max(L, 1), -1, np.int32)
        self._chk(lib().orbf_fuse_search(self._h, int(slot), _p(Rcw), _p(tcw), _p(camera), opt(kp_x) if slot < 0 else None, opt(kp_y) if slot < 0 else None,
                                         opt(u_right) if slot < 0 else None, opt(desc) if slot < 0 else None, n, opt(lm_pos), opt(lm_desc), opt(lm_valid), L,
                                         C.c_float(radius), int(th_low), _p(best), _p(dist)), "fuse_search")
        return best[:L], dist[:L]

    def bow_match(self, words1, off1, idx1, desc1, words2, off2, idx2, desc2, nn_ratio=0.6, th_low=50):
        """Matcher::BoWMatch (Features/matcher.cpp:145-209) on flattened DBoW3 feature vectors: DMatch array in the reference's order."""
        w1 = np.ascontiguousarray(words1, np.int32); o1 = np.ascontiguousarray(off1, np.int32); i1 = np.ascontiguousarray(idx1, np.int32)
        w2 = np.ascontiguousarray(words2, np.int32); o2 = np.ascontiguousarray(off2, np.int32); i2 = np.ascontiguousarray(idx2, np.int32)
        d1 = np.ascontiguousarray(desc1, np.uint8).reshape(-1, 32); d2 = np.ascontiguousarray(desc2, np.uint8).reshape(-1, 32)
        opt = lambda a: _p(a) if len(a) else None
        out = np.zeros(max(len(i1), 1), DMATCH_DT); n = C.c_int32(0)
        self._chk(lib().orbf_bow_match(self._h, opt(w1), _p(o1), opt(i1), len(w1), opt(d1), len(d1), opt(w2), _p(o2), opt(i2), len(w2), opt(d2), len(d2),
                                       C.c_float(nn_ratio), int(th_low), _p(out), len(out), C.byref(n)), "bow_match")
        return out[:n.value].copy()

    def match_pairs(self, pairs, ratio, cross_check=False):
        pairs = np.ascontiguousarray(pairs, np.int32).reshape(-1, 2)
        self._chk(lib().orbf_match_pairs(self._h, _p(pairs), len(pairs), C.c_float(ratio), int(cross_check)), "match_pairs")
        return len(pairs)

    def download_matches(self, pair):
        out = np.zeros(self.K, DMATCH_DT); n = C.c_int32(0)
        self._chk(lib().orbf_download_matches(self._h, pair, _p(out), self.K, C.byref(n)), "download_matches")
        return out[:n.value].copy()

    def download_knn(self, pair):
        o = [np.zeros(self.K, np.int32) for _ in range(4)]; n = C.c_int32(0)
        self._chk(lib().orbf_download_knn(self._h, pair, _p(o[0]), _p(o[1]), _p(o[2]), _p(o[3]), self.K, C.byref(n)), "download_knn")
        return tuple(a[:n.value].copy() for a in o)

    def match_counts(self, npairs):
        out = np.zeros(npairs, np.int32)
        self._chk(lib().orbf_match_counts(self._h, npairs, _p(out)), "match_counts")
        return out

    # ---- RANSAC / Kabsch ----
    def ransac_iterate(self, src_xyz, dst_xyz, m12, sample_table=None, want_table=True, **kw):
        """want_table=False leaves sample_table_out NULL: the library then draws the sample table lazily (first rows in
        ransac_prepare, the rest only if the loop gets that far), as the batched paths do; the returned table is all -1."""
        cfg = default_ransac_config(**kw)
        src = np.ascontiguousarray(src_xyz, np.float32); dst = np.ascontiguousarray(dst_xyz, np.float32)
        m12 = np.ascontiguousarray(m12, DMATCH_DT)
        res = RansacResult()
        inl = np.zeros(max(len(m12), 1), DMATCH_DT); hyp = np.zeros(cfg.iterations, HYP_DT)
        good = np.zeros(max(len(m12), 1), DMATCH_DT); tab_out = np.full((cfg.iterations, cfg.sample_size), -1, np.int32)
        tab = None if sample_table is None else np.ascontiguousarray(sample_table, np.int32)
        self._chk(lib().orbf_ransac_iterate(self._h, C.byref(cfg), _p(src), len(src), _p(dst), len(dst), _p(m12), len(m12), _p(tab),
                                            _p(inl), len(inl), C.byref(res), _p(hyp), _p(good), _p(tab_out) if want_table else None), "ransac_iterate")
        return dict(ok=bool(res.ok), rmse=float(res.rmse), T12=np.array(res.T12, np.float32).reshape(4, 4),
                    inliers=inl[:res.n_inliers].copy(), n_good=res.n_good, real_iters=res.real_iters, valid_iters=res.valid_iters,
                    used_identity=bool(res.used_identity), depth_cov=float(res.depth_cov_used), hyp=hyp,
                    good_sorted=good[:res.n_good].copy(), sample_table=tab_out)

    def odometry_compute(self, src_xyz, dst_xyz, m12, pose1=None, **kw):
        """Odometry::Compute, RANSAC strategy, for one frame pair in one call (Odometry/odometry.cpp:44-90): the ransac_iterate result plus
        cloud_src / cloud_tgt ([n, 4] pcl::PointXYZ records) and pose2 = T12 * pose1."""
        cfg = default_ransac_config(**kw)
        src = np.ascontiguousarray(src_xyz, np.float32); dst = np.ascontiguousarray(dst_xyz, np.float32)
        m12 = np.ascontiguousarray(m12, DMATCH_DT)
        res = RansacResult(); inl = np.zeros(max(len(m12), 1), DMATCH_DT)
        cs = np.zeros((max(len(m12), 1), 4), np.float32); ct = np.zeros_like(cs); nc = C.c_int32(0)
        p1 = None if pose1 is None else np.ascontiguousarray(pose1, np.float32).reshape(16)
        p2 = np.zeros(16, np.float32)
        self._chk(lib().orbf_odometry_compute(self._h, C.byref(cfg), _p(src), len(src), _p(dst), len(dst), _p(m12), len(m12), _p(inl), len(inl), C.byref(res),
                                              _p(cs), _p(ct), len(m12), C.byref(nc), _p(p1) if p1 is not None else None, _p(p2)), "odometry_compute")
        return dict(ok=bool(res.ok), rmse=float(res.rmse), T12=np.array(res.T12, np.float32).reshape(4, 4), inliers=inl[:res.n_inliers].copy(),
                    n_good=res.n_good, depth_cov=float(res.depth_cov_used), cloud_src=cs[:nc.value].copy(), cloud_tgt=ct[:nc.value].copy(),
                    pose2=p2.reshape(4, 4))

    def ransac_pairs(self, npairs, **kw):
        cfg = default_ransac_config(**kw)
        self._chk(lib().orbf_ransac_pairs(self._h, npairs, C.byref(cfg)), "ransac_pairs")

    def ransac_probe_depth_cov(self, npairs, **kw):
        """The depth covariance (quirk Q7) the pairs last matched would latch on a fresh process, or -1."""
        cfg = default_ransac_config(**kw); cov = C.c_double(-1.0)
        self._chk(lib().orbf_ransac_probe_depth_cov(self._h, npairs, C.byref(cfg), C.byref(cov)), "ransac_probe_depth_cov")
        return cov.value

    def download_ransac(self, pair):
        res = RansacResult(); inl = np.zeros(self.K, DMATCH_DT)
        self._chk(lib().orbf_download_ransac(self._h, pair, C.byref(res), _p(inl), self.K), "download_ransac")
        return dict(ok=bool(res.ok), rmse=float(res.rmse), T12=np.array(res.T12, np.float32).reshape(4, 4),
                    inliers=inl[:res.n_inliers].copy(), n_good=res.n_good, real_iters=res.real_iters, valid_iters=res.valid_iters,
                    used_identity=bool(res.used_identity), depth_cov=float(res.depth_cov_used))

    def download_ransac_summary(self, npairs):
        out = np.zeros(npairs, RANSAC_RESULT_DT)
        self._chk(lib().orbf_download_ransac_summary(self._h, npairs, _p(out)), "download_ransac_summary")
        return out

    def download_ransac_clouds(self, pair):
        """Ransac::mpSourceCloud / mpTargetCloud of a pair last solved: two [n, 4] float32 arrays (x, y, z, 1 — pcl::PointXYZ records)."""
        a = np.zeros((self.K, 4), np.float32); b = np.zeros((self.K, 4), np.float32); n = C.c_int32(0)
        self._chk(lib().orbf_download_ransac_clouds(self._h, pair, _p(a), _p(b), self.K, C.byref(n)), "download_ransac_clouds")
        return a[:n.value].copy(), b[:n.value].copy()

    def ransac_clouds_device(self, pair0, npairs):
        """Device-resident clouds of pairs [pair0, pair0 + npairs): (src pointer, tgt pointer, counts pointer, points per pair)."""
        ps = C.c_void_p(); pt = C.c_void_p(); pc = C.c_void_p(); k = C.c_int32(0)
        self._chk(lib().orbf_ransac_clouds(self._h, pair0, npairs, C.byref(ps), C.byref(pt), C.byref(pc), C.byref(k)), "ransac_clouds")
        return ps.value, pt.value, pc.value, k.value

    def kabsch(self, A, B):
        A = np.ascontiguousarray(A, np.float32).reshape(-1, 3); B = np.ascontiguousarray(B, np.float32).reshape(-1, 3)
        T = np.zeros(16, np.float32)
        self._chk(lib().orbf_kabsch(self._h, _p(A) if len(A) else None, _p(B) if len(B) else None, len(A), _p(T)), "kabsch")
        return T.reshape(4, 4)

    # ---- adaptive-threshold FAST detector (Extractor ADAPTIVE mode, BASELINE config 4) ----
    def adaptive_detect(self, frames, thresh, **kw):
        """frames [n, H, W] u8; thresh [grid*grid] float64 state, updated in place.  Returns (list of keypoint arrays,
        cell_thresh [n, cells], cell_found [n, cells])."""
        cfg = AdaptiveConfig()
        lib().orbf_default_adaptive_config(C.byref(cfg))
        for k, v in kw.items():
            setattr(cfg, k, v)
        frames = np.ascontiguousarray(frames, np.uint8)
        n, h, w = frames.shape
        cells = cfg.grid * cfg.grid
        assert thresh.dtype == np.float64 and len(thresh) == cells
        cap = cfg.max_per_cell * cells
        out = np.zeros((n, cap), KEYPOINT_DT); counts = np.zeros(n, np.int32)
        used = np.zeros((n, cells), np.int32); found = np.zeros((n, cells), np.int32)
        self._chk(lib().orbf_adaptive_detect(self._h, C.byref(cfg), _p(frames), n, w, C.c_int64(w * h), _p(thresh), _p(out), _p(counts), cap,
                                             _p(used), _p(found)), "adaptive_detect")
        return [out[i, :counts[i]].copy() for i in range(n)], used, found

    def selftest_sincosf_device(self, lo_bits, hi_bits):
        """Device restatement of glibc sinf / cosf over the floats with bit patterns [lo, hi]: [n, 2] (sin, cos)."""
        out = np.zeros((int(hi_bits) - int(lo_bits) + 1, 2), np.float32)
        self._chk(lib().orbf_selftest_sincosf_device(self._h, C.c_uint32(lo_bits), C.c_uint32(hi_bits), _p(out)), "selftest_sincosf_device")
        return out

    def extract_adapted(self, frames, thresh, depths=None, slot0=0, **kw):
        """BASELINE config 4, 8-level variant: ORB extraction with per-region adapted iniThFAST, frames in order; `thresh` [grid*grid]
        float64 is updated in place.  Returns (region_th [n, g2], region_found [n, g2]); results live in the frame slots."""
        cfg = AdaptiveConfig()
        lib().orbf_default_adaptive_config(C.byref(cfg))
        for k, v in kw.items():
            setattr(cfg, k, v)
        frames = np.ascontiguousarray(frames, np.uint8)
        n, h, w = frames.shape
        g2 = cfg.grid * cfg.grid
        assert thresh.dtype == np.float64 and len(thresh) == g2
        used = np.zeros((n, g2), np.int32); found = np.zeros((n, g2), np.int32)
        self._chk(lib().orbf_extract_adapted(self._h, slot0, n, _p(frames), C.c_int64(w), C.c_int64(w * h), _p(depths), C.c_int64(w), C.c_int64(w * h),
                                             C.byref(cfg), _p(thresh), _p(used), _p(found)), "extract_adapted")
        return used, found

    def extract_adapted_videos(self, videos, thresh, slot0=0, **kw):
        """V independent videos advancing together: videos [V, T, H, W] u8, thresh [V, grid*grid] float64 (updated in place).  Frame t of
        video v lands in slot slot0 + t * V + v.  Returns (region_th [T, V, g2], region_found [T, V, g2])."""
        cfg = AdaptiveConfig()
        lib().orbf_default_adaptive_config(C.byref(cfg))
        for k, v in kw.items():
            setattr(cfg, k, v)
        videos = np.ascontiguousarray(videos, np.uint8)
        V, T, h, w = videos.shape
        g2 = cfg.grid * cfg.grid
        assert thresh.dtype == np.float64 and thresh.shape == (V, g2) and thresh.flags.c_contiguous
        used = np.zeros((T, V, g2), np.int32); found = np.zeros((T, V, g2), np.int32)
        self._chk(lib().orbf_extract_adapted_videos(self._h, slot0, V, T, _p(videos), C.c_int64(w), C.c_int64(w * h), C.c_int64(T * w * h), None, C.c_int64(w),
                                                    C.c_int64(w * h), C.c_int64(T * w * h), C.byref(cfg), _p(thresh), _p(used), _p(found)), "extract_adapted_videos")
        return used, found

    # ---- keyframe store ----
    def kfdb_reserve(self, n):
        self._chk(lib().orbf_kfdb_reserve(self._h, n), "kfdb_reserve")

    def kfdb_add_from_slot(self, kf, slot):
        self._chk(lib().orbf_kfdb_add_from_slot(self._h, kf, slot), "kfdb_add_from_slot")

    def kfdb_add_host(self, kf, desc):
        desc = np.ascontiguousarray(desc, np.uint8)
        self._chk(lib().orbf_kfdb_add_host(self._h, kf, _p(desc), len(desc)), "kfdb_add_host")

    def kfdb_device_buffers(self):
        d = C.c_void_p(None); cnt = C.c_void_p(None); rows = C.c_int32(0); n = C.c_int32(0)
        self._chk(lib().orbf_kfdb_device_buffers(self._h, C.byref(d), C.byref(cnt), C.byref(rows), C.byref(n)), "kfdb_device_buffers")
        return d.value, cnt.value, rows.value, n.value

    def kfdb_attach_device(self, d_desc_ptr, d_counts_ptr, n_kf):
        self._chk(lib().orbf_kfdb_attach_device(self._h, C.c_void_p(d_desc_ptr) if d_desc_ptr else None,
                                                C.c_void_p(d_counts_ptr) if d_counts_ptr else None, n_kf), "kfdb_attach_device")

    def kfdb_match_slot(self, slot, kf0, nkf, ratio):
        """Device-resident query (frame slot) against keyframes [kf0, kf0 + nkf); asynchronous."""
        self._chk(lib().orbf_kfdb_match_slot(self._h, slot, kf0, nkf, C.c_float(ratio)), "kfdb_match_slot")

    def kfdb_results(self, nkf, nq, tables=True):
        o = [np.zeros((nkf, nq), np.int32) for _ in range(4)] if tables else [None] * 4
        surv = np.zeros(nkf, np.int32)
        self._chk(lib().orbf_kfdb_results(self._h, nkf, nq, _p(o[0]), _p(o[1]), _p(o[2]), _p(o[3]), _p(surv)), "kfdb_results")
        return o[0], o[1], o[2], o[3], surv

    # ---- multi-GPU keyframe store ----
    def comm_init(self, id128, nranks, rank):
        id128 = np.ascontiguousarray(id128, np.uint8); assert id128.size == 128
        self._chk(lib().orbf_comm_init(self._h, _p(id128), nranks, rank), "comm_init")

    def comm_destroy(self):
        self._chk(lib().orbf_comm_destroy(self._h), "comm_destroy")

    def kfdb_allgather(self):
        n = C.c_int32(0)
        self._chk(lib().orbf_kfdb_allgather(self._h, C.byref(n)), "kfdb_allgather")
        return n.value

    def kfdb_ipc_handles(self):
        a = np.zeros(64, np.uint8); b = np.zeros(64, np.uint8)
        self._chk(lib().orbf_kfdb_ipc_handles(self._h, _p(a), _p(b)), "kfdb_ipc_handles")
        return a, b

    def kfdb_attach_peers(self, desc_handles, count_handles, nranks, rank, kf_per_rank):
        dh = np.ascontiguousarray(desc_handles, np.uint8).reshape(nranks, 64); ch = np.ascontiguousarray(count_handles, np.uint8).reshape(nranks, 64)
        self._chk(lib().orbf_kfdb_attach_peers(self._h, _p(dh), _p(ch), nranks, rank, kf_per_rank), "kfdb_attach_peers")

    def kfdb_detach_peers(self):
        self._chk(lib().orbf_kfdb_detach_peers(self._h), "kfdb_detach_peers")

    def kfdb_survivors(self, q, kf0, nkf, ratio):
        """Number of ratio-test survivors of the query against each keyframe (no top-2 tables downloaded)."""
        q = np.ascontiguousarray(q, np.uint8)
        surv = np.zeros(nkf, np.int32)
        self._chk(lib().orbf_kfdb_match(self._h, _p(q), len(q), kf0, nkf, C.c_float(ratio), None, None, None, None, _p(surv)), "kfdb_match")
        return surv

    def kfdb_match(self, q, kf0, nkf, ratio):
        q = np.ascontiguousarray(q, np.uint8)
        o = [np.zeros((nkf, len(q)), np.int32) for _ in range(4)]; surv = np.zeros(nkf, np.int32)
        self._chk(lib().orbf_kfdb_match(self._h, _p(q), len(q), kf0, nkf, C.c_float(ratio), _p(o[0]), _p(o[1]), _p(o[2]), _p(o[3]),
                                        _p(surv)), "kfdb_match")
        return o[0], o[1], o[2], o[3], surv


# ---- host-side selftest hooks of csrc/replay.h (no GPU needed) ----
def selftest_introsort(m):
    m = np.ascontiguousarray(m, DMATCH_DT).copy()
    rc = lib().orbf_selftest_introsort(_p(m), len(m))
    if rc:
        raise OrbfError(rc, "selftest_introsort")
    return m


def selftest_glibc_rand(seed, n):
    out = np.zeros(n, np.int32)
    rc = lib().orbf_selftest_glibc_rand(C.c_uint32(seed), n, _p(out))
    if rc:
        raise OrbfError(rc, "selftest_glibc_rand")
    return out


def selftest_sample_table(seed, M, iterations=200, sample_size=4):
    out = np.zeros((iterations, sample_size), np.int32)
    rc = lib().orbf_selftest_sample_table(C.c_uint32(seed), M, iterations, sample_size, _p(out))
    if rc:
        raise OrbfError(rc, "selftest_sample_table")
    return out


def selftest_sincosf(lo_bits, hi_bits, want_values=False):
    """Host restatement of glibc sinf / cosf (csrc/glibc_sincosf.h) over the floats with bit patterns [lo, hi]:
    (number of inputs where it differs from this machine's libm, values [n, 2] (sin, cos) or None)."""
    n = int(hi_bits) - int(lo_bits) + 1
    out = np.zeros((n, 2), np.float32) if want_values else None
    nd = C.c_int64(0)
    rc = lib().orbf_selftest_sincosf(C.c_uint32(lo_bits), C.c_uint32(hi_bits), _p(out), C.byref(nd))
    if rc:
        raise OrbfError(rc, "selftest_sincosf")
    return nd.value, out


def comm_unique_id():
    """128-byte NCCL unique id (rank 0 draws it, the application broadcasts it, every rank passes it to Context.comm_init)."""
    out = np.zeros(128, np.uint8)
    rc = lib().orbf_comm_unique_id(_p(out))
    if rc:
        raise OrbfError(rc, "comm_unique_id")
    return out
