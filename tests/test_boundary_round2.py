"""Boundary completions of round 2 through the C ABI:
  * Ransac::mpSourceCloud / mpTargetCloud (Odometry/ransac.cpp:163-189) of the pairs last solved — batched (slot-based) and standalone;
  * the per-keypoint tail of Frame::ExtractFeatures (Core/frame.cpp:138-164) for keypoints that did not come out of the ORB extractor,
    with and without the reference's FR1 distortion;
CPU part: the library exports the new entry points and the oracle's numpy restatements behave as the reference lines say."""
import ctypes as C

import numpy as np
import pytest

import synth

FR1_DIST = (0.262383, -0.953104, -0.005358, 0.002628, 1.163314)     # k1, k2, p1, p2, k3 (Utils/common.h:40-44)


def test_new_entry_points_are_exported(ob):
    L = ob.lib()
    for name in ("orbf_ransac_clouds", "orbf_download_ransac_clouds", "orbf_unproject_keypoints"):
        assert hasattr(L, name), name


def test_oracle_clouds_follow_the_reference_rules(orc):
    rng = np.random.default_rng(3)
    src = rng.uniform(0.5, 4, (50, 3)).astype(np.float32); dst = rng.uniform(0.5, 4, (60, 3)).astype(np.float32)
    src[5, 2] = 0; dst[7, 2] = np.nan; src[9, 2] = -1
    m = np.zeros(40, orc.DMATCH_DT); m["queryIdx"] = np.arange(40); m["trainIdx"] = np.arange(40)[::-1]
    s, t = orc.ransac_clouds(src, dst, m)
    bad = {5, 9} | {int(np.where(m["trainIdx"] == 7)[0][0])}
    keep = [i for i in range(40) if i not in bad]
    assert len(s) == len(t) == 37 and np.array_equal(s[:, :3], src[m["queryIdx"][keep]]) and np.all(s[:, 3] == 1) and np.array_equal(t[:, :3], dst[m["trainIdx"][keep]])
    s, t = orc.ransac_clouds(src, dst, m, check_depth=False)
    assert len(s) == 40
    s, t = orc.ransac_clouds(src, dst, m[:19])                   # fewer than mMinInlierTh: cleared and left empty (ransac.cpp:163-167)
    assert len(s) == 0 and len(t) == 0


def test_oracle_keyframe_matcher_filter(orc):
    rng = np.random.default_rng(5)
    q = rng.integers(0, 256, (40, 32), dtype=np.uint8)
    t = np.concatenate([q[::-1].copy(), rng.integers(0, 256, (10, 32), dtype=np.uint8)])       # exact copies: every query survives the ratio test
    base = orc.knn_match(q, t, 0.8, False)
    assert len(base) == 40
    lm1 = np.arange(1, 41); lm1[3] = 0
    lm2 = np.zeros(50, np.int64); lm2[39 - 10] = 99                                               # train row of query 10 already holds a landmark
    got = orc.knn_match_keyframe(q, t, 0.8, lm1, lambda p: p == 21, lm2)
    assert sorted(set(range(40)) - set(got["queryIdx"].tolist())) == [3, 10, 20]
    assert lm2[39 - 0] == 1 and lm2[39 - 10] == 99


@pytest.mark.gpu
def test_ransac_clouds_batched_and_standalone(ob, orc, texture):
    n = 4
    frames = np.stack([synth.make_frame(texture, 30 + i) for i in range(n)])
    depths = np.stack([synth.make_depth(30 + i) for i in range(n)])
    depths[1, ::3] = 0                                          # plenty of matches fail the depth check on pairs 0 and 1
    ctx = ob.Context(max_frames=n)
    try:
        ctx.track_sequence(frames, depths, 0.8, cross_check=True, seed=42); ctx.synchronize()
        host = []
        for i in range(n):
            k, d = orc.extract(frames[i]); host.append((d, orc.unproject(k, depths[i])[0]))
        for p in range(n - 1):
            m = orc.knn_match(host[p][0], host[p + 1][0], 0.8, True)
            ws, wt = orc.ransac_clouds(host[p][1], host[p + 1][1], m)
            gs, gt = ctx.download_ransac_clouds(p)
            assert len(gs) == len(ws) and np.array_equal(gs, ws) and np.array_equal(gt, wt), f"pair {p}"
            if p < 2:
                assert len(ws) < len(m)
        ps, pt, pc, k = ctx.ransac_clouds_device(0, n - 1)
        assert ps and pt and pc and k == ctx.K
        # standalone call = pair 0; fewer than min_inlier_th matches leave the clouds empty
        m = orc.knn_match(host[2][0], host[3][0], 0.8, True)
        ctx.ransac_iterate(host[2][1], host[3][1], m, seed=7)
        gs, gt = ctx.download_ransac_clouds(0)
        ws, wt = orc.ransac_clouds(host[2][1], host[3][1], m)
        assert np.array_equal(gs, ws) and np.array_equal(gt, wt) and len(ws) > 20
        ctx.ransac_iterate(host[2][1], host[3][1], m[:10], seed=7)
        gs, gt = ctx.download_ransac_clouds(0)
        assert len(gs) == 0 and len(gt) == 0
    finally:
        ctx.close()


@pytest.mark.gpu
def test_odometry_compute_is_iterate_plus_clouds_plus_composition(ob, orc):
    """orbf_odometry_compute (one call, one synchronisation) against the oracle's Ransac::Iterate, the clouds rule and the composition
    rule of Odometry::Compute (odometry.cpp:82-84) — and against the three separate calls it replaces."""
    ctx = ob.Context(max_frames=2)
    try:
        rng = np.random.default_rng(2)
        ang = 0.3
        pose1 = np.eye(4, dtype=np.float32)
        pose1[:3, :3] = np.array([[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]], np.float32)
        pose1[:3, 3] = rng.normal(0, 1, 3).astype(np.float32)
        # easy pairs end inside the first two hypothesis waves; the hard ones (tens to all 200 iterations) take the call's second phase
        for seed, outl in ((42, 0.3), (7, 0.6), (9, 1.0), (21, 0.6), (22, 0.75), (23, 1.0), (24, 0.3)):
            src, dst, m, _, _ = synth.rigid_pairs(seed=seed, outlier_frac=outl)
            src = src.copy(); src[m["queryIdx"][::9], 2] = np.nan            # some matches fail the depth check: clouds shorter than m
            g = ctx.odometry_compute(src, dst, m, pose1=pose1, seed=seed)
            r = orc.ransac_iterate(src, dst, m, seed=seed)
            if outl >= 0.75:
                assert r["real_iters"] > 8
            assert g["ok"] == r["ok"] and g["inliers"].tobytes() == r["inliers"].tobytes()
            assert np.array_equal(g["T12"], r["T12"]) and g["rmse"] == r["rmse"] and g["depth_cov"] == r["depth_cov"]
            ws, wt = orc.ransac_clouds(src, dst, m)
            assert np.array_equal(g["cloud_src"], ws) and np.array_equal(g["cloud_tgt"], wt) and len(ws) < len(m)
            assert np.array_equal(g["pose2"], orc.compose_trajectory(r["T12"][None], pose1)[1])
            s = ctx.ransac_iterate(src, dst, m, seed=seed)                     # the separate calls
            cs, ct = ctx.download_ransac_clouds(0)
            assert s["inliers"].tobytes() == g["inliers"].tobytes() and np.array_equal(cs, g["cloud_src"]) and np.array_equal(ct, g["cloud_tgt"])
        g = ctx.odometry_compute(src, dst, m[:10], pose1=pose1)                # fewer than min_inlier_th matches: T12 = I, pose2 = pose1
        assert not g["ok"] and len(g["cloud_src"]) == 0 and np.array_equal(g["pose2"], pose1) and np.array_equal(g["T12"], np.eye(4, dtype=np.float32))
        g = ctx.odometry_compute(src, dst, m, seed=3)                          # pose1 = NULL: identity
        assert np.array_equal(g["pose2"], orc.compose_trajectory(g["T12"][None])[1])
    finally:
        ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("dist", [None, FR1_DIST])
def test_unproject_keypoints_matches_oracle(ob, orc, texture, dist):
    frame = synth.make_frame(texture, 11); depth = synth.make_depth(11)
    th = np.full(9, 20.0)
    kps = orc.adaptive_detect(frame, th, retain_best=1000)[0]
    kw = {} if dist is None else dict(k1=dist[0], k2=dist[1], p1=dist[2], p2=dist[3], k3=dist[4])
    ctx = ob.Context(max_frames=1, **kw)
    try:
        xyz, ur, un = ctx.unproject_keypoints(kps, depth)
        wx, wu = orc.unproject(kps, depth, dist=dist)
        assert np.array_equal(xyz, wx) and np.array_equal(ur, wu)
        xy = np.stack([kps["x"], kps["y"]], 1).astype(np.float32)
        want_un = xy if dist is None else orc.undistort_points(xy, 517.3, 516.5, 318.6, 255.3, dist)
        assert np.array_equal(un, want_un)
        xyz0, ur0, _ = ctx.unproject_keypoints(kps, None)
        assert not xyz0.any() and np.all(ur0 == -1)
        assert (xyz[:, 2] > 0).sum() > 0.8 * len(kps)
    finally:
        ctx.close()
