"""The oracle's restatement of the per-frame front end against THE REFERENCE'S OWN SOURCE with its REAL Core classes:
oracle/_ref/libframe_ref.so is Core/frame.cpp + keyframe.cpp + landmark.cpp + map.cpp + Features/extractor.cpp + orbextractor.cpp
(+ matcher.cpp) compiled verbatim.  Frame::Frame (BGR -> gray, depth scale, the distortion vector of Utils/common.h) and
Frame::ExtractFeatures (rows a-9, a-10, 8f-2: Extractor::Extract -> ORBextractor, UndistortKeyPoints, the depth gather at the truncated
distorted position, mvuRight, the unprojection of the undistorted point, ComputeImageBounds) must reproduce the oracle's keypoints,
descriptors, mvKeysUn, mvKeys3Dc, mvuRight bit for bit; Landmark::ComputeDistinctiveDescriptors (8f-1) the oracle's choice.

CPU-only.  Skipped where neither the reference checkout nor a prebuilt oracle/_ref exists."""
import numpy as np
import pytest

import synth

FR1 = dict(fx=517.3, fy=516.5, cx=318.6, cy=255.3)
FR1_DIST = np.array([0.262383, -0.953104, -0.005358, 0.002628, 1.163314], np.float32)     # k1, k2, p1, p2, k3 (Utils/common.h:40-44)


@pytest.fixture(scope="module")
def ref():
    from oracle import ref as r
    if not r.available():
        pytest.skip("oracle/_ref not built and /root/reference absent")
    return r


def _bgr(gray_like, seed):
    """A colour frame whose channels differ, so that the BGR -> gray weights matter."""
    rng = np.random.default_rng(seed)
    g = gray_like.astype(np.int16)
    b = np.clip(g + rng.integers(-20, 21, g.shape), 0, 255); r = np.clip(g - rng.integers(-20, 21, g.shape), 0, 255)
    return np.stack([b, g, r], -1).astype(np.uint8)


@pytest.mark.parametrize("i", [0, 3, 8])
def test_frame_extract_features_identical_to_reference_source(ref, orc, texture, i):
    bgr = _bgr(synth.make_frame(texture, i), i)
    depth = synth.make_depth(i)
    depth[::7, ::5] = 0                                                  # holes: no 3D point, mvuRight stays -1
    r = ref.frame_extract(bgr, depth)
    gray = orc.bgr2gray(bgr)
    assert np.array_equal(r["gray"], gray)                               # Frame::Frame: cvtColor(BGR2GRAY)
    k, d = orc.extract(gray)
    assert len(k) == len(r["kps"]) > 900
    assert r["kps"].tobytes() == k.tobytes() and np.array_equal(r["desc"], d)
    xy = np.stack([k["x"], k["y"]], 1).astype(np.float32)
    un = orc.undistort_points(xy, FR1["fx"], FR1["fy"], FR1["cx"], FR1["cy"], FR1_DIST)
    assert np.array_equal(r["xy_un"], un)                                # Frame::UndistortKeyPoints
    assert np.abs(un - xy).max() > 0.5                                   # the FR1 distortion does move the points
    xyz, ur = orc.unproject(k, depth, dist=FR1_DIST)
    assert np.array_equal(r["xyz"], xyz) and np.array_equal(r["uright"], ur)
    assert (ur < 0).any() and (ur >= 0).sum() > 700                      # both branches of `if (z > 0)`
    # Frame::ComputeImageBounds: the undistorted image corners
    corners = np.array([[0, 0], [640, 0], [0, 480], [640, 480]], np.float32)
    c = orc.undistort_points(corners, FR1["fx"], FR1["fy"], FR1["cx"], FR1["cy"], FR1_DIST)
    exp = np.array([min(c[0, 0], c[2, 0]), max(c[1, 0], c[3, 0]), min(c[0, 1], c[1, 1]), max(c[2, 1], c[3, 1])], np.float32)
    assert np.array_equal(r["bounds"], exp)


def test_compute_distinctive_descriptors(ref, orc):
    rng = np.random.default_rng(4)
    nobs = rng.integers(0, 20, 300); nobs[:3] = (0, 1, 2)
    offs = np.concatenate([[0], np.cumsum(nobs)]).astype(np.int32)
    base = rng.integers(0, 256, (len(nobs), 32), dtype=np.uint8)
    desc = np.repeat(base, nobs, axis=0)
    flips = rng.integers(0, 256, (len(desc), 12))
    for c in range(flips.shape[1]):                                       # noisy copies: medians tie often, the first smallest must win
        desc[np.arange(len(desc)), flips[:, c] // 8] ^= (1 << (flips[:, c] % 8)).astype(np.uint8)
    out, has = ref.distinctive_descriptors(desc, offs)
    best, _ = orc.distinctive_descriptors(desc, offs)
    for l in range(len(nobs)):
        if nobs[l] == 0:
            assert not has[l] and best[l] == -1
        else:
            assert has[l] and np.array_equal(out[l], desc[offs[l] + best[l]]), f"landmark {l}"


def test_distinctive_descriptors_skip_bad_keyframes(ref, orc):
    """landmark.cpp:235-238: observations whose keyframe isBad() are left out before the medians are taken."""
    rng = np.random.default_rng(9)
    nobs = rng.integers(3, 15, 80)
    offs = np.concatenate([[0], np.cumsum(nobs)]).astype(np.int32)
    desc = rng.integers(0, 256, (int(offs[-1]), 32), dtype=np.uint8)
    bad = (rng.random(len(desc)) < 0.3).astype(np.uint8)
    bad[offs[5]:offs[6]] = 1                                              # a landmark whose every observer is bad keeps no descriptor
    out, has = ref.distinctive_descriptors(desc, offs, bad)
    keep = bad == 0
    offs2 = np.concatenate([[0], np.cumsum([int(keep[offs[l]:offs[l + 1]].sum()) for l in range(len(nobs))])]).astype(np.int32)
    kept = desc[keep]
    best, _ = orc.distinctive_descriptors(kept, offs2)
    assert not has[5]
    for l in range(len(nobs)):
        if offs2[l + 1] > offs2[l]:
            assert has[l] and np.array_equal(out[l], kept[offs2[l] + best[l]]), f"landmark {l}"


def test_odometry_compute_on_real_frames(ref, orc):
    """Odometry(RANSAC).Compute from Odometry/odometry.cpp + ransac.cpp on the reference's real Frame objects: Ransac::Iterate, the
    composition rule T12 * pF1->GetPose() through cv::Mat (SetPose / GetPose of frame.cpp), SetInlier(m.trainIdx) — against the oracle's
    ransac_iterate + compose_trajectory."""
    cov = ref.frame_depth_covariance(2.0)
    rng = np.random.default_rng(1)
    ang = 0.2
    pose1 = np.eye(4, dtype=np.float32)
    pose1[:3, :3] = np.array([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]], np.float32)
    pose1[:3, 3] = rng.normal(0, 1, 3).astype(np.float32)
    for seed, outl in ((42, 0.3), (7, 0.6), (9, 1.0)):
        src, dst, m, _, _ = synth.rigid_pairs(seed=seed, outlier_frac=outl)
        r = ref.odometry_compute(src, dst, m, pose1, seed=seed)
        o = orc.ransac_iterate(src, dst, m, seed=seed, depth_cov=cov)
        assert r["inliers"].tobytes() == o["inliers"].tobytes() and np.array_equal(r["T12"], o["T12"]) and r["rmse"] == o["rmse"]
        assert np.array_equal(r["pose2"], orc.compose_trajectory(o["T12"][None], pose1)[1])
        exp = np.ones(len(dst), bool); exp[o["inliers"]["trainIdx"]] = False
        assert np.array_equal(r["outlier2"], exp)


def test_knn_match_on_real_frames_and_landmarks(ref, orc):
    """Matcher::KnnMatch(Frame&, Frame&) compiled against the reference's real Core classes: the kNN-2 + ratio survivors of the oracle, then
    the reference's landmark rules replayed here in query order — F1's feature must hold a landmark and not be an outlier; F2's slot must
    not hold a landmark with Observations() > 0; an accepted match moves the landmark into the slot and flags it as outlier."""
    rng = np.random.default_rng(11)
    q, t = synth.descriptor_sets(n=600, seed=5)[:2]
    obs1 = rng.integers(0, 4, len(q)).astype(np.int32); obs1[rng.random(len(q)) < 0.2] = -1
    out1 = (rng.random(len(q)) < 0.1).astype(np.uint8)
    obs2 = np.full(len(t), -1, np.int32); pick = rng.random(len(t)) < 0.15; obs2[pick] = rng.integers(0, 3, int(pick.sum()))
    got, slot2, outl2 = ref.real_knn_match_frames(q, t, 0.8, obs1, out1, obs2)
    base = orc.knn_match(q, t, 0.8, False)
    slot_obs = obs2.copy(); exp = []; exp_slot = np.full(len(t), -1, np.int32); exp_out = np.zeros(len(t), bool)
    for m in base:
        i1, i2 = int(m["queryIdx"]), int(m["trainIdx"])
        if obs1[i1] < 0 or out1[i1]:
            continue
        if slot_obs[i2] > 0:
            continue
        slot_obs[i2] = obs1[i1]; exp_slot[i2] = i1; exp_out[i2] = True
        exp.append(m)
    exp = np.array(exp, base.dtype)
    assert got.tobytes() == exp.tobytes() and 100 < len(exp) < len(base)
    assert np.array_equal(slot2, exp_slot) and np.array_equal(outl2, exp_out)
