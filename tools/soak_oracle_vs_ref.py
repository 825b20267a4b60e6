#!/usr/bin/env python3
"""Randomised soak of the CPU oracle against the reference's OWN sources (oracle/_ref: orbextractor.cpp, matcher.cpp, ransac.cpp compiled
verbatim) — authoring container only (needs the reference checkout).  The committed tests (tests/test_oracle_vs_ref*.py) compare fixed
cases; this draws random ones for a given number of seconds and reports every disagreement:

  extract  random geometry (160..900 x 120..700), nfeatures 50..2500, 1..9 levels, scale 1.1..2.0, ini / min thresholds, six image
           kinds (noise, blurred noise, blocks, low contrast = th-7 fallback, checkerboards = ties, the bench's synthetic motion frames):
           keypoints incl. order and descriptors byte for byte
  ransac   random rigid problems (30..1500 points, 0..100 % outliers, depth holes / NaN, tie-heavy distances), iterations 1..400,
           sample size 3..6, inlier threshold, Mahalanobis threshold, depth check on / off, both Iterate forms: ok, inlier list incl.
           order, T12, rmse bit for bit
  match    random descriptor sets (1..1500 query rows, 2..1500 train rows, near-duplicates, heavy ties), ratio 0.5..1.0: the DMatch list byte for byte

  adaptive random clips of 3..8 frames (160..900 x 120..700, the same image kinds, contrast changes in mid-clip) through
           Extractor(FAST, ., ADAPTIVE): the nine controllers' state, the keypoint count and the multiset of responses frame by frame
           (keypoints themselves only up to quirk Q15: ties at the std::nth_element cut)

  8f       random scenes of the tests' generators (tests/test_projection_match.py, tests/test_fuse_bow.py) through ProjectionMatch,
           BoWMatch and the search of Fuse (matcher.cpp:90-313) with random sizes, radii, ratios and thresholds: indices / DMatch lists

  frame    random 640x480 colour frames (the calibration compiled into the reference) with random depth planes (holes, far / near
           ranges) through the reference's real Frame::Frame + Frame::ExtractFeatures (Core/frame.cpp:18-45,135-170,286-313): gray plane,
           keypoints, descriptors, undistorted keypoints, 3D points, mvuRight bit for bit

  cv2      what oracle/_ref CANNOT see — the OpenCV arithmetic its stand-in headers share with the oracle — against the real library
           (cv2 4.x): cv2.resize(INTER_LINEAR) at random source / destination sizes, cv2.GaussianBlur(7x7, 2, REFLECT_101), cv2.FAST with
           NMS on random ROIs and thresholds (positions, order, responses), cv2.fastAtan2, BFMatcher.knnMatch(k = 2) order under ties,
           cv2.cvtColor(BGR2GRAY), cv2.undistortPoints (5 iterations, P = K) with random coefficients, and the whole extraction against
           the cv2-driven restatement tests/cv2_oracle.py

  replay   csrc/replay.h — the std::sort and glibc rand() restatements the RANSAC KERNELS run, host build of the same code inside
           liborbfront_b200.so — against the real libstdc++ std::sort on DMatch (0..3000 elements, tie-heavy / sorted / reversed /
           median-of-3-killer inputs: the permutation of equal keys must be the library's) and against libc srand / rand sample tables

  linalg   the arithmetic of the libraries that are in NEITHER the image NOR oracle/_ref (PCL TransformationFromCorrespondences, Eigen
           JacobiSVD / LLT: restated from their published algorithms, f32 like the reference) against numpy / LAPACK in f64, to
           tolerance: 3x3 SVD (random, rank-deficient, tiny, diagonal), the weighted rigid transform on noisy correspondences (weights
           1 / (z1 z2), ransac.cpp:308), Kabsch incl. reflections, the Mahalanobis distance of ErrorFunction2 (ransac.cpp:350-414).
           A case fails above 2e-5 (relative; 1e-9 for the f64 distance); the JSON line also carries the largest deviation seen

  landmark random observation sets (0..40 noisy copies of a descriptor per landmark, medians tie often, random bad keyframes) through
           the reference's real Landmark::ComputeDistinctiveDescriptors over KeyFrame / Landmark / Map objects (Core/landmark.cpp:219-273)

  odometry Odometry(RANSAC).Compute (odometry.cpp:44-90) and Matcher::KnnMatch(Frame&, Frame&) with its landmark rules (matcher.cpp:55-88) on
           the reference's real Frame / Landmark objects: inliers, T12, rmse, the composed pose T12 * pose1 through cv::Mat, the inlier
           flags; matches, the landmarks handed to the second frame and its outlier flags

  python tools/soak_oracle_vs_ref.py <extract|ransac|match|adaptive|8f|frame|landmark|odometry|cv2|replay|linalg> <seed> <seconds>      -> one JSON line
"""
import json
import sys
import time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "tests")]
import synth
from oracle import oracle as orc, ref

mode, seed, seconds = sys.argv[1], int(sys.argv[2]), float(sys.argv[3])
orc.build(); assert ref.available()
rng = np.random.default_rng(seed)
t0 = time.time(); n = 0; bad = []


def image(w, h):
    import cv2
    kind = int(rng.integers(0, 6))
    if kind == 0:
        return kind, rng.integers(0, 256, (h, w)).astype(np.uint8)
    if kind == 1:
        img = cv2.GaussianBlur(rng.integers(0, 256, (h, w)).astype(np.uint8), (0, 0), float(rng.uniform(0.7, 4)))
        return kind, cv2.normalize(img, None, 0, 255, cv2.NORM_MINMAX)
    if kind == 2:
        img = np.full((h, w), int(rng.integers(0, 256)), np.uint8)
        for _ in range(int(rng.integers(1, 200))):
            x, y, s = int(rng.integers(0, w - 8)), int(rng.integers(0, h - 8)), int(rng.integers(2, 9))
            img[y:y + s, x:x + s] = int(rng.integers(0, 256))
        return kind, img
    if kind == 3:
        base = cv2.GaussianBlur(rng.integers(0, 256, (h, w)).astype(np.uint8), (0, 0), 2.0).astype(np.float32)
        return kind, np.clip(128 + (base - base.mean()) * float(rng.uniform(0.2, 1.5)), 0, 255).astype(np.uint8)
    if kind == 4:
        s = int(rng.integers(3, 20)); yy, xx = np.mgrid[0:h, 0:w]
        return kind, (((yy // s + xx // s) % 2) * int(rng.integers(30, 255))).astype(np.uint8)
    tex = synth.make_texture(int(rng.integers(0, 1000)), h, w)
    return kind, synth.make_frame(tex, int(rng.integers(0, 30)), w, h, seed=int(rng.integers(0, 99)))


if mode == "ransac":
    cov = ref.depth_covariance(1.0)            # quirk Q7: latch the reference's process-wide covariance first, hand the same to the oracle
while time.time() - t0 < seconds:
    if mode == "extract":
        w, h = int(rng.integers(160, 900)), int(rng.integers(120, 700))
        p = dict(nfeatures=int(rng.integers(50, 2500)), nlevels=int(rng.integers(1, 10)), scale_factor=float(rng.choice([1.1, 1.2, 1.3, 1.5, 2.0])),
                 ini_th=int(rng.choice([20, 20, 20, 12, 40])), min_th=int(rng.choice([7, 7, 5, 3])))
        kind, img = image(w, h)
        try:
            k, d = orc.extract(img, **p)
        except Exception:
            continue                              # geometry too small for this pyramid (ORBF-style error in the oracle; the reference asserts)
        rk, rd = ref.extract(img, **p)
        same = k.tobytes() == rk.tobytes() and np.array_equal(d, rd)
        what = dict(w=w, h=h, kind=kind, **p)
    elif mode == "ransac":
        n_pts = int(rng.integers(30, 1500)); m = int(rng.integers(0, n_pts + 1))
        src, dst, mm, _, _ = synth.rigid_pairs(m=max(m, 12), seed=int(rng.integers(0, 10 ** 6)), outlier_frac=float(rng.choice([0.0, 0.05, 0.3, 0.5, 0.7, 0.9, 1.0])),
                                               n_pts=max(n_pts, 12))
        mm = mm[:m]
        tie = int(rng.integers(0, 4))
        if tie == 0: mm["distance"] = rng.integers(0, 8, len(mm))
        elif tie == 1: mm["distance"] = rng.integers(0, 100, len(mm))
        elif tie == 2: mm["distance"] = 5
        p = dict(iterations=int(rng.choice([200, 200, 50, 1, 400])), min_inlier_th=int(rng.choice([20, 20, 5, 60])), max_mahal=float(rng.choice([3.0, 3.0, 1.0, 10.0])),
                 sample_size=int(rng.choice([4, 4, 3, 6])), check_depth=bool(rng.integers(0, 4) > 0), seed=int(rng.integers(0, 2 ** 31)))
        r = ref.ransac_iterate(src, dst, mm, member_form=bool(rng.integers(0, 2)) and p["check_depth"], **p)
        o = orc.ransac_iterate(src, dst, mm, depth_cov=cov, **p)
        same = (r["ok"] == o["ok"] and r["n_good"] == o["n_good"] and r["inliers"].tobytes() == o["inliers"].tobytes()
                and np.array_equal(r["T12"], o["T12"], equal_nan=True) and (r["rmse"] == o["rmse"] or (np.isnan(r["rmse"]) and np.isnan(o["rmse"]))))
        what = dict(n_pts=n_pts, m=len(mm), tie=tie, **p)
    elif mode == "adaptive":
        from collections import Counter
        w, h = int(rng.integers(160, 900)), int(rng.integers(120, 700))
        nfr = int(rng.integers(3, 9)); kind0, base = image(w, h)
        ex = ref.AdaptiveExtractor(); th = np.full(9, 20.0); same = True
        for f in range(nfr):
            if kind0 == 5 or rng.integers(0, 3) == 0:
                _, img = image(w, h)
            else:
                img = np.roll(base, (int(rng.integers(-5, 6)), int(rng.integers(-5, 6))), (0, 1))
            if rng.integers(0, 3) == 0:
                img = (img.astype(np.float32) * float(rng.uniform(0.1, 1.0)) + float(rng.uniform(0, 100))).clip(0, 255).astype(np.uint8)
            rk, rth = ex.extract(img)
            ok, _, _ = orc.adaptive_detect(img, th, retain_best=1000)
            same &= bool(np.array_equal(rth, th)) and len(rk) == len(ok) and Counter(rk["response"].tolist()) == Counter(ok["response"].tolist())
            if not same:
                break
        ex.close()
        what = dict(w=w, h=h, kind=kind0, frames=nfr, failed_at=f)
    elif mode == "8f":
        from test_fuse_bow import CAM, _bow_scene, _fuse_scene
        from test_projection_match import _scene as _proj_scene
        which = int(rng.integers(0, 3)); sd = int(rng.integers(0, 10 ** 6))
        if which == 0:
            nfeat, nlm = int(rng.integers(20, 900)), int(rng.integers(5, 900))
            sc = _proj_scene(sd, n_feat=nfeat, n_lm=nlm, crowded=bool(rng.integers(0, 2)))
            kw = dict(radius=float(rng.choice([3.0, 8.0, 15.0])), nn_ratio=float(rng.choice([0.6, 0.8, 0.95])), th_high=float(rng.choice([60.0, 100.0, 256.0])))
            taken = sc[8] if rng.integers(0, 2) else None
            b_r, n_r = ref.projection_match(*sc[:8], feat_taken=taken, **kw); b_o, n_o = orc.projection_match(*sc[:8], feat_taken=taken, **kw)
            same = bool(np.array_equal(b_r, b_o)) and n_r == n_o
            what = dict(fn="projection", seed=sd, n_feat=nfeat, n_lm=nlm, **kw)
        elif which == 1:
            n1, n2, nw = int(rng.integers(10, 900)), int(rng.integers(10, 900)), int(rng.integers(2, 300))
            sc = _bow_scene(sd, n1=n1, n2=n2, n_words=nw)
            kw = dict(nn_ratio=float(rng.choice([0.6, 0.75, 0.9])), th_low=float(rng.choice([50.0, 80.0, 30.0])))
            same = ref.bow_match(*sc, **kw).tobytes() == orc.bow_match(*sc, **kw).tobytes()
            what = dict(fn="bow", seed=sd, n1=n1, n2=n2, n_words=nw, **kw)
        else:
            nfeat, nlm, radius = int(rng.integers(20, 900)), int(rng.integers(5, 800)), float(rng.choice([3.0, 8.0]))
            R, t, kp_x, kp_y, u_right, desc, pw, lm_desc, valid = _fuse_scene(sd, n_feat=nfeat, n_lm=nlm, radius=radius)
            state = valid.copy(); inval = np.nonzero(valid == 0)[0]
            state[inval] = np.array([0, 2, 3], np.uint8)[np.arange(len(inval)) % 3]
            th = float(rng.choice([50.0, 30.0, 100.0]))
            b_r, nf = ref.fuse(R, t, CAM[5:9], kp_x, kp_y, u_right, desc, pw, lm_desc, state, radius=radius, th_low=th)
            b_o, _ = orc.fuse_search(R, t, CAM, kp_x, kp_y, u_right, desc, pw, lm_desc, valid, radius=radius, th_low=th)
            same = bool(np.array_equal(b_r, b_o)) and nf == int((b_o >= 0).sum())
            what = dict(fn="fuse", seed=sd, n_feat=nfeat, n_lm=nlm, radius=radius, th_low=th)
    elif mode == "frame":
        from test_oracle_vs_ref_frame import FR1, FR1_DIST
        kind, g = image(640, 480)
        gi = g.astype(np.int16)
        bgr = np.stack([np.clip(gi + rng.integers(-40, 41, g.shape), 0, 255), gi, np.clip(gi - rng.integers(-40, 41, g.shape), 0, 255)], -1).astype(np.uint8)
        depth = rng.integers(int(rng.choice([0, 2000])), int(rng.choice([8000, 40000, 65536])), (480, 640)).astype(np.uint16)
        if rng.integers(0, 2):
            depth[rng.random((480, 640)) < float(rng.uniform(0.01, 0.6))] = 0
        r = ref.frame_extract(bgr, depth)
        gray = orc.bgr2gray(bgr)
        k, d = orc.extract(gray)
        xy = np.stack([k["x"], k["y"]], 1).astype(np.float32).reshape(-1, 2)
        un = orc.undistort_points(xy, FR1["fx"], FR1["fy"], FR1["cx"], FR1["cy"], FR1_DIST) if len(k) else xy
        xyz, ur = orc.unproject(k, depth, dist=FR1_DIST)
        same = (bool(np.array_equal(r["gray"], gray)) and r["kps"].tobytes() == k.tobytes() and bool(np.array_equal(r["desc"], d)) and bool(np.array_equal(r["xy_un"], un))
                and bool(np.array_equal(r["xyz"], xyz)) and bool(np.array_equal(r["uright"], ur)))
        what = dict(kind=kind, nk=len(k))
    elif mode == "cv2":
        import cv2
        import cv2_oracle as co
        which = int(rng.integers(0, 8))
        if which == 0:
            sw, sh = int(rng.integers(8, 900)), int(rng.integers(8, 700)); f = float(rng.uniform(1.05, 2.2))
            dw, dh = max(int(round(sw / f)), 2), max(int(round(sh / f)), 2)
            _, img = image(max(sw, 16), max(sh, 16)); img = np.ascontiguousarray(img[:sh, :sw])
            same = bool(np.array_equal(orc.resize_linear(img, dw, dh), cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR)))
            what = dict(fn="resize", src=(sw, sh), dst=(dw, dh))
        elif which == 1:
            w, h = int(rng.integers(4, 900)), int(rng.integers(4, 700))
            _, img = image(max(w, 16), max(h, 16)); img = np.ascontiguousarray(img[:h, :w])
            same = bool(np.array_equal(orc.gaussian_blur7(img), cv2.GaussianBlur(img.copy(), (7, 7), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101)))
            what = dict(fn="blur", size=(w, h))
        elif which == 2:
            w, h = int(rng.integers(7, 300)), int(rng.integers(7, 300)); th = int(rng.integers(1, 80))
            _, img = image(max(w, 16), max(h, 16)); roi = np.ascontiguousarray(img[:h, :w])
            kps = cv2.FastFeatureDetector_create(th, True).detect(roi)
            want = [(int(k.pt[0]), int(k.pt[1]), int(k.response)) for k in kps]
            got = [(int(c["x"]), int(c["y"]), int(c["score"])) for c in orc.fast_roi(roi, th)]
            same = got == want
            what = dict(fn="fast", size=(w, h), th=th)
        elif which == 3:
            ys = rng.integers(-400000, 400000, 3000).astype(np.float32); xs = rng.integers(-400000, 400000, 3000).astype(np.float32)
            ys[:20] = 0; xs[20:40] = 0; ys[40:60] = xs[40:60]; ys[60:80] = -xs[60:80]
            same = all(np.float32(cv2.fastAtan2(float(y), float(x))) == np.float32(orc.fast_atan2(y, x)) for y, x in zip(ys, xs))
            what = dict(fn="fastAtan2")
        elif which == 4:
            nq, nt = int(rng.integers(1, 400)), int(rng.integers(2, 400))
            q = rng.integers(0, 256, (nq, 32)).astype(np.uint8); t = rng.integers(0, 256, (nt, 32)).astype(np.uint8)
            if rng.integers(0, 2):
                t = t[rng.integers(0, min(nt, 5), nt)]
            same = all(bool(np.array_equal(g, r)) for g, r in zip(orc.knn2(q, t), co.knn2(q, t)))
            what = dict(fn="knn2", nq=nq, nt=nt)
        elif which == 5:
            w, h = int(rng.integers(1, 700)), int(rng.integers(1, 500))
            bgr = rng.integers(0, 256, (h, w, 3)).astype(np.uint8)
            same = bool(np.array_equal(orc.bgr2gray(bgr), cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY)))
            what = dict(fn="bgr2gray", size=(w, h))
        elif which == 6:
            fx, fy, cx, cy = float(rng.uniform(300, 900)), float(rng.uniform(300, 900)), float(rng.uniform(200, 700)), float(rng.uniform(150, 500))
            dist = np.array([rng.uniform(-0.4, 0.4), rng.uniform(-1.0, 1.0), rng.uniform(-0.01, 0.01), rng.uniform(-0.01, 0.01), rng.choice([0.0, rng.uniform(-1.2, 1.2)])], np.float32)
            pts = np.stack([rng.uniform(-5, 2 * cx + 5, 500), rng.uniform(-5, 2 * cy + 5, 500)], 1).astype(np.float32)
            K = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], np.float32)
            want = cv2.undistortPoints(pts.reshape(-1, 1, 2), K, dist, None, K).reshape(-1, 2)
            same = bool(np.array_equal(orc.undistort_points(pts, np.float32(fx), np.float32(fy), np.float32(cx), np.float32(cy), dist), want))
            what = dict(fn="undistortPoints", K=(fx, fy, cx, cy), dist=dist.tolist())
        else:
            w, h = int(rng.integers(200, 700)), int(rng.integers(160, 520))
            p = dict(nfeatures=int(rng.integers(100, 1500)), nlevels=int(rng.integers(1, 9)), scale_factor=float(rng.choice([1.2, 1.2, 1.3, 1.5])))
            kind, img = image(w, h)
            try:
                k, d = orc.extract(img, **p)
            except Exception:
                continue
            try:
                k2, d2, _ = co.extract(img, orc.pattern(), **p)
            except (OverflowError, ZeroDivisionError):
                continue                          # a level narrower than one 30-px cell: width / nCols with nCols = 0 (orbextractor.cpp:683-686), Python raises
            same = bool(np.array_equal(k, np.array(k2, dtype=orc.KEYPOINT_DT))) and bool(np.array_equal(d, d2))
            what = dict(fn="extract_via_cv2", w=w, h=h, kind=kind, **p)
    elif mode == "replay":
        import bench
        ob = bench.load_pkg()
        if rng.integers(0, 3):
            m = int(rng.integers(0, 3000)); kind = int(rng.integers(0, 6))
            d = np.zeros(m, ob.DMATCH_DT); d["queryIdx"] = np.arange(m); d["trainIdx"] = rng.permutation(m) if m else 0
            if kind == 0: d["distance"] = rng.integers(0, 64, m)
            elif kind == 1: d["distance"] = rng.integers(0, 3, m)
            elif kind == 2: d["distance"] = np.sort(rng.integers(0, 256, m))[::-1]
            elif kind == 3: d["distance"] = rng.random(m)
            elif kind == 4: d["distance"] = np.sort(rng.integers(0, 40, m))
            elif m >= 2:
                k = m // 2; a = np.zeros(m, np.float32)
                for i in range(k):
                    a[i] = i + 1 if i % 2 == 0 else k + i + (1 if k % 2 else 0)
                    a[k + i] = 2 * (i + 1)
                d["distance"] = a
            same = ob.selftest_introsort(d).tobytes() == orc.std_sort_dmatch(d).tobytes()
            what = dict(fn="std::sort", m=m, kind=kind)
        else:
            sd, M = int(rng.integers(0, 2 ** 32)), int(rng.integers(1, 2000))
            same = bool(np.array_equal(ob.selftest_sample_table(sd, M), orc.sample_table_libc(sd, M))) and bool(np.array_equal(ob.selftest_glibc_rand(sd, 400), orc.libc_rand_sequence(sd, 400)))
            what = dict(fn="rand", seed=sd, M=M)
    elif mode == "linalg":
        which = int(rng.integers(0, 4)); dev = 0.0
        if which == 0:
            A = rng.normal(size=(3, 3)).astype(np.float32); k = int(rng.integers(0, 5))
            if k == 1: A[:, 2] = A[:, 0] * 2
            elif k == 2: A *= np.float32(10.0 ** int(rng.integers(-8, 6)))
            elif k == 3: A = np.diag(rng.normal(size=3)).astype(np.float32)
            U, S, V = orc.svd3(A); scale = max(float(np.abs(A).max()), 1e-30)
            sn = np.linalg.svd(A.astype(np.float64), compute_uv=False)
            dev = max(float(np.abs(U @ np.diag(S) @ V.T - A).max()) / scale, float(np.abs(U.T @ U - np.eye(3)).max()), float(np.abs(V.T @ V - np.eye(3)).max()),
                      float(np.abs(S - sn).max()) / max(float(sn.max()), 1e-30))
            same = dev < 2e-5 and bool(S[0] >= S[1] >= S[2] >= 0)
            what = dict(fn="svd3", kind=k, dev=dev)
        elif which in (1, 2):
            npts = int(rng.integers(4, 600))
            P = np.empty((npts, 3)); P[:, 2] = rng.uniform(0.5, 5.0, npts); P[:, 0] = rng.uniform(-0.6, 0.6, npts) * P[:, 2]; P[:, 1] = rng.uniform(-0.45, 0.45, npts) * P[:, 2]
            ax = rng.normal(size=3); ax /= np.linalg.norm(ax); ang = float(rng.uniform(0, 0.6))
            Kx = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
            R = np.eye(3) + np.sin(ang) * Kx + (1 - np.cos(ang)) * Kx @ Kx; t = rng.normal(0, 0.3, 3)
            Q = P @ R.T + t + rng.normal(0, float(rng.choice([0.0, 0.002, 0.02])), (npts, 3))
            Q[:, 2] = np.maximum(Q[:, 2], 0.05)
            P32, Q32 = P.astype(np.float32), Q.astype(np.float32)
            Pd, Qd = P32.astype(np.float64), Q32.astype(np.float64)
            w = 1.0 / (Pd[:, 2] * Qd[:, 2]) if which == 1 else np.ones(npts)
            m1 = (w[:, None] * Pd).sum(0) / w.sum(); m2 = (w[:, None] * Qd).sum(0) / w.sum()
            Cm = ((w[:, None] * (Qd - m2)).T @ (Pd - m1)) / w.sum()
            Uu, Ss, Vt = np.linalg.svd(Cm)
            Rr = Uu @ np.diag([1, 1, np.sign(np.linalg.det(Uu) * np.linalg.det(Vt))]) @ Vt
            want = np.eye(4); want[:3, :3] = Rr; want[:3, 3] = m2 - Rr @ m1
            T = orc.weighted_transform(P32, Q32) if which == 1 else orc.kabsch(P32, Q32)
            cond = Ss[1] / max(Ss[0], 1e-30)
            if cond < 1e-3:
                continue                          # a (nearly) collinear set: the rotation about the line is not determined
            dev = float(np.abs(T - want).max())
            same = dev < 2e-5 * max(1.0, float(np.abs(Qd).max())) / min(1.0, cond * 10)
            what = dict(fn="weighted_transform" if which == 1 else "kabsch", n=npts, dev=dev, cond=float(cond))
        else:
            p = np.array([rng.uniform(-2, 2), rng.uniform(-2, 2), rng.uniform(0.4, 6)], np.float32)
            ax = rng.normal(size=3); ax /= np.linalg.norm(ax); ang = float(rng.uniform(0, 0.5))
            Kx = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
            T = np.eye(4, dtype=np.float32); T[:3, :3] = (np.eye(3) + np.sin(ang) * Kx + (1 - np.cos(ang)) * Kx @ Kx).astype(np.float32); T[:3, 3] = rng.normal(0, 0.2, 3).astype(np.float32)
            Td = T.astype(np.float64)
            q = (Td[:3, :3] @ p.astype(np.float64) + Td[:3, 3] + rng.normal(0, float(rng.choice([0.001, 0.01, 0.05])), 3)).astype(np.float32)
            if not q[2] > 0.05:
                continue
            cz = float(rng.choice([1.3e-3, 1e-4, 2.5e-2]))
            cx = (3 * np.tan(58.0 / 180.0 * np.pi / 640)) ** 2; cy = (3 * np.tan(45.0 / 180.0 * np.pi / 480)) ** 2
            mu = Td[:3, :3] @ p.astype(np.float64) + Td[:3, 3]; dl = mu - q.astype(np.float64)
            got = orc.mahalanobis2(p, q, T, cz)
            s1 = max(cx * float(p[2]), cy * float(p[2]), cz); s2 = max(cx * float(q[2]), cy * float(q[2]), cz)
            if dl @ dl > 2 * (s1 + s2) * 0.98:
                continue                          # at or beyond the shortcut reject (ransac.cpp:375-381): the function returns DBL_MAX there
            Sg = Td[:3, :3].T @ np.diag([cx * p[2], cy * p[2], cz]) @ Td[:3, :3] + np.diag([cx * q[2], cy * q[2], cz])
            want = float(dl @ np.linalg.solve(Sg, dl))
            if got == np.finfo(np.float64).max:
                continue
            dev = abs(got - want) / max(1.0, abs(want))
            same = dev <= 1e-9
            what = dict(fn="mahalanobis2", dev=dev)
        worst = globals().setdefault("worst", {}); worst[what["fn"]] = max(worst.get(what["fn"], 0.0), dev)
    elif mode == "landmark":
        nl = int(rng.integers(1, 200)); nobs = rng.integers(0, int(rng.choice([4, 12, 41])), nl)
        offs = np.concatenate([[0], np.cumsum(nobs)]).astype(np.int32)
        base = rng.integers(0, 256, (nl, 32), dtype=np.uint8); desc = np.repeat(base, nobs, axis=0)
        flips = rng.integers(0, 256, (len(desc), int(rng.integers(0, 40))))
        for c in range(flips.shape[1]):
            desc[np.arange(len(desc)), flips[:, c] // 8] ^= (1 << (flips[:, c] % 8)).astype(np.uint8)
        badkf = (rng.random(len(desc)) < float(rng.choice([0.0, 0.2, 0.7]))).astype(np.uint8)
        out, has = ref.distinctive_descriptors(desc, offs, badkf)
        keep = badkf == 0
        offs2 = np.concatenate([[0], np.cumsum([int(keep[offs[l]:offs[l + 1]].sum()) for l in range(nl)])]).astype(np.int32)
        kept = np.ascontiguousarray(desc[keep]).reshape(-1, 32)
        best, _ = orc.distinctive_descriptors(kept, offs2) if len(kept) else (np.full(nl, -1, np.int32), None)
        same = True
        for l in range(nl):
            if offs2[l + 1] > offs2[l]:
                same &= bool(has[l]) and bool(np.array_equal(out[l], kept[offs2[l] + best[l]]))
            else:
                same &= not bool(has[l])
        what = dict(n_landmarks=nl, rows=int(len(desc)))
    elif mode == "odometry":
        if "cov_frame" not in globals():
            cov_frame = ref.frame_depth_covariance(2.0)           # quirk Q7 in the frame library's own process-wide static
        if rng.integers(0, 2):
            ang = float(rng.uniform(-1, 1)); pose1 = np.eye(4, dtype=np.float32)
            pose1[:3, :3] = np.array([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]], np.float32)
            pose1[:3, 3] = rng.normal(0, 1, 3).astype(np.float32)
            n_pts = int(rng.integers(30, 1200)); m = int(rng.integers(1, n_pts + 1)); sd = int(rng.integers(0, 2 ** 31))
            src, dst, mm, _, _ = synth.rigid_pairs(m=max(m, 12), seed=int(rng.integers(0, 10 ** 6)), outlier_frac=float(rng.choice([0.0, 0.3, 0.6, 0.9, 1.0])), n_pts=max(n_pts, 12))
            if rng.integers(0, 2):
                mm["distance"] = rng.integers(0, 30, len(mm))
            r = ref.odometry_compute(src, dst, mm, pose1, seed=sd)
            o = orc.ransac_iterate(src, dst, mm, seed=sd, depth_cov=cov_frame)
            exp = np.ones(len(dst), bool); exp[o["inliers"]["trainIdx"]] = False
            same = (r["inliers"].tobytes() == o["inliers"].tobytes() and bool(np.array_equal(r["T12"], o["T12"], equal_nan=True)) and (r["rmse"] == o["rmse"] or (np.isnan(r["rmse"]) and np.isnan(o["rmse"])))
                    and bool(np.array_equal(r["outlier2"], exp)))
            if o["ok"]:
                same = same and bool(np.array_equal(r["pose2"], orc.compose_trajectory(o["T12"][None], pose1)[1]))
            what = dict(fn="Odometry::Compute", n_pts=n_pts, m=len(mm), seed=sd, ok=bool(o["ok"]))
        else:
            nq, nt = int(rng.integers(1, 900)), int(rng.integers(2, 900))
            q = rng.integers(0, 256, (nq, 32)).astype(np.uint8); t = rng.integers(0, 256, (nt, 32)).astype(np.uint8)
            for j in range(0, nt, 2):
                t[j] = q[int(rng.integers(0, nq))]; t[j, int(rng.integers(0, 32))] ^= 1 << int(rng.integers(0, 8))
            ratio = float(rng.choice([0.6, 0.8, 0.95]))
            obs1 = rng.integers(0, 4, nq).astype(np.int32); obs1[rng.random(nq) < 0.2] = -1
            out1 = (rng.random(nq) < 0.1).astype(np.uint8)
            obs2 = np.full(nt, -1, np.int32); pick = rng.random(nt) < 0.15; obs2[pick] = rng.integers(0, 3, int(pick.sum()))
            got, slot2, outl2 = ref.real_knn_match_frames(q, t, ratio, obs1, out1, obs2)
            base = orc.knn_match(q, t, ratio, False)
            slot_obs = obs2.copy(); exp = []; exp_slot = np.full(nt, -1, np.int32); exp_out = np.zeros(nt, bool)
            for mt in base:
                i1, i2 = int(mt["queryIdx"]), int(mt["trainIdx"])
                if obs1[i1] < 0 or out1[i1] or slot_obs[i2] > 0:
                    continue
                slot_obs[i2] = obs1[i1]; exp_slot[i2] = i1; exp_out[i2] = True
                exp.append(mt)
            exp = np.array(exp, base.dtype)
            same = got.tobytes() == exp.tobytes() and bool(np.array_equal(slot2, exp_slot)) and bool(np.array_equal(outl2, exp_out))
            what = dict(fn="KnnMatch(Frame&, Frame&)", nq=nq, nt=nt, ratio=ratio)
    else:
        nq, nt = int(rng.integers(1, 1500)), int(rng.integers(2, 1500))      # nt = 1: the reference reads matchesKnn[i][1] of a one-element vector (matcher.cpp:64), undefined
        q = rng.integers(0, 256, (nq, 32)).astype(np.uint8); t = rng.integers(0, 256, (nt, 32)).astype(np.uint8)
        kind = int(rng.integers(0, 3))
        if kind == 1:                             # near-duplicates of the query rows in the train set
            for j in range(0, nt, 2):
                t[j] = q[int(rng.integers(0, nq))]; t[j, int(rng.integers(0, 32))] ^= 1 << int(rng.integers(0, 8))
        elif kind == 2:                           # few distinct rows: ties everywhere
            t = t[rng.integers(0, min(nt, 4), nt)]; q = q[rng.integers(0, min(nq, 6), nq)]
        ratio = float(rng.choice([0.5, 0.6, 0.8, 0.9, 1.0]))
        a = ref.knn_match_frames(q, t, ratio); b = orc.knn_match(q, t, ratio, False)
        same = a.tobytes() == b.tobytes()
        what = dict(nq=nq, nt=nt, kind=kind, ratio=ratio)
    n += 1
    if not same:
        bad.append(what)
res = {"mode": mode, "seed": seed, "seconds": round(time.time() - t0, 1), "cases": n, "mismatches": len(bad), "first_mismatches": bad[:5]}
if mode == "linalg":
    res["largest_deviation"] = globals().get("worst", {})
print(json.dumps(res))
