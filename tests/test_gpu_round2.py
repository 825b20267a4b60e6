"""GPU parity cases added in round 2 (VERDICT r1 "next round" item 1):
  * the depth covariance (quirk Q7) is latched by the first pair that REACHES scoring, not by pair slot 0;
  * the whole chain at BASELINE config 4's geometry (1280x720, 2000 keypoints): extract -> kNN-2 (~2000 x 2000, cross-check on and
    off) -> RANSAC on a 2000-keypoint context, against the oracle;
  * Frame::UndistortKeyPoints feeding mvuRight / mvKeys3Dc with the reference's own FR1 distortion (Utils/common.h:40-44);
  * the 512-frame pipelined sequence compared with the oracle on every distinct pair kind and on 72 pairs byte for byte.
"""
import numpy as np
import pytest

import synth

pytestmark = pytest.mark.gpu

FR1_DIST = (0.262383, -0.953104, -0.005358, 0.002628, 1.163314)     # k1, k2, p1, p2, k3 (Utils/common.h:40-44)


def _oracle_chain(orc, frames, depths, ratio, cross, seed, dist=None, **okw):
    """The reference's per-frame loop on the oracle, the covariance carried from pair to pair as the process-static would be."""
    host = []
    for f, d in zip(frames, depths):
        k, desc = orc.extract(f, **okw)
        xyz, ur = orc.unproject(k, d, dist=dist)
        host.append((k, desc, xyz, ur))
    cov = -1.0; out = []
    for p in range(len(frames) - 1):
        m = orc.knn_match(host[p][1], host[p + 1][1], ratio, cross) if len(host[p][1]) and len(host[p + 1][1]) >= 2 else np.zeros(0, orc.DMATCH_DT)
        r = orc.ransac_iterate(host[p][2], host[p + 1][2], m, seed=seed + p, depth_cov=cov)
        cov = r["depth_cov"]
        out.append((m, r))
    return host, out


def _assert_pair(ctx, p, m, r):
    assert ctx.download_matches(p).tobytes() == m.tobytes(), f"pair {p}: matches"
    g = ctx.download_ransac(p)
    assert g["ok"] == r["ok"] and g["n_good"] == r["n_good"], f"pair {p}: ok / n_good"
    assert g["inliers"].tobytes() == r["inliers"].tobytes(), f"pair {p}: inliers"
    assert np.abs(g["T12"] - r["T12"]).max() <= 1e-5 and g["rmse"] == r["rmse"], f"pair {p}: pose"     # north-star tolerance 1e-5
    assert g["depth_cov"] == r["depth_cov"], f"pair {p}: depth covariance {g['depth_cov']} vs {r['depth_cov']}"
    assert g["real_iters"] == r["real_iters"] and g["valid_iters"] == r["valid_iters"]


@pytest.mark.parametrize("mode", ["flat_first_frame", "no_depth_first_frame", "two_dead_pairs"])
@pytest.mark.parametrize("pipelined", [False, True])
def test_depth_covariance_latched_by_first_scoring_pair(ob, orc, texture, mode, pipelined):
    n = 6
    frames = np.stack([synth.make_frame(texture, 60 + i) for i in range(n)])
    depths = np.stack([synth.make_depth(60 + i) for i in range(n)])
    if mode == "flat_first_frame":
        frames[0] = 128                         # no keypoints: pair 0 has no matches (ransac.cpp:165 returns before any scoring)
    elif mode == "no_depth_first_frame":
        depths[0] = 0                           # every match of pair 0 is dropped by the depth check (ransac.cpp:175-189, :191)
    else:
        frames[0] = 128; depths[2] = 0          # pairs 0, 1 and 2 never score; pair 3 latches
    _, ref = _oracle_chain(orc, frames, depths, 0.8, True, 42)
    assert not ref[0][1]["ok"] and ref[0][1]["depth_cov"] < 0, "pair 0 must not reach scoring in this case"
    ctx = ob.Context(max_frames=n, pipeline_chunk=2 if pipelined else -1, pipeline_streams=2)
    try:
        if pipelined:
            ctx.track_sequence(frames, depths, 0.8, cross_check=True, seed=42)
        else:
            ctx.extract_batch(frames, depths)
            ctx.match_pairs(np.array([[i, i + 1] for i in range(n - 1)], np.int32), 0.8, cross_check=True)
            ctx.ransac_pairs(n - 1, seed=42)
        summ = ctx.download_ransac_summary(n - 1)
        covs = [float(r["depth_cov"]) for _, r in ref]
        latched = [c for c in covs if c >= 0][0]
        # the device value is per call (every pair of one call reports the covariance the call ended up with)
        assert np.all(summ["depth_cov_used"] == latched)
        for p, (m, r) in enumerate(ref):
            g = ctx.download_ransac(p)
            assert ctx.download_matches(p).tobytes() == m.tobytes()
            assert g["ok"] == r["ok"] and g["inliers"].tobytes() == r["inliers"].tobytes(), f"pair {p}"
            assert np.abs(g["T12"] - r["T12"]).max() <= 1e-5
        assert sum(int(r["ok"]) for _, r in ref) >= 2, "the later pairs must still be solved"
    finally:
        ctx.close()


def test_standalone_ransac_does_not_touch_the_latched_covariance(ob, orc, texture):
    n = 3
    frames = np.stack([synth.make_frame(texture, 10 + i) for i in range(n)])
    depths = np.stack([synth.make_depth(10 + i) for i in range(n)])
    src, dst, matches, _, _ = synth.rigid_pairs(m=300, seed=3, outlier_frac=0.3, n_pts=400)
    ctx = ob.Context(max_frames=n)
    try:
        ctx.track_sequence(frames, depths, 0.8, cross_check=True, seed=1)
        before = ctx.download_ransac_summary(n - 1)["depth_cov_used"].copy()
        a = ctx.ransac_iterate(src * 3.0, dst * 3.0, matches, seed=9)                    # latches its own value (other depths)
        b = ctx.ransac_iterate(src, dst, matches, seed=9, depth_cov=1.0e-4)              # explicit value
        ra = orc.ransac_iterate(src * 3.0, dst * 3.0, matches, seed=9)
        rb = orc.ransac_iterate(src, dst, matches, seed=9, depth_cov=1.0e-4)
        assert a["depth_cov"] == ra["depth_cov"] and a["inliers"].tobytes() == ra["inliers"].tobytes()
        assert b["depth_cov"] == rb["depth_cov"] and b["inliers"].tobytes() == rb["inliers"].tobytes()
        assert a["depth_cov"] != before[0]
        ctx.track_sequence(frames, depths, 0.8, cross_check=True, seed=1)                # still scores with the value latched first
        assert np.array_equal(ctx.download_ransac_summary(n - 1)["depth_cov_used"], before)
        # download with inliers == NULL and cap == 0 is a summary read, not a capacity error
        import ctypes as C
        res = ob.RansacResult()
        assert ob.lib().orbf_download_ransac(ctx._h, 0, C.byref(res), None, 0) == 0 and res.n_inliers > 0
    finally:
        ctx.close()


def test_config4_geometry_full_chain_1280x720_2000kp(ob, orc):
    """BASELINE config 4 geometry end to end: K = 2048 rows per frame, kNN-2 at ~2000 x 2000, RANSAC with the hypothesis kernel's
    shared memory at its largest (nfeatures 1961..2024 used to fail to launch: static + dynamic > 48 KB without the opt-in)."""
    w, h, n = 1280, 720, 3
    tex = synth.make_texture(5, h, w)
    frames = np.stack([synth.make_frame(tex, i, w, h, seed=5) for i in range(n)])
    depths = np.stack([synth.make_depth(i, w, h, seed=5) for i in range(n)])
    okw = dict(nfeatures=2000)
    fx, fy, cx, cy = synth.FX * w / 640.0, synth.FY * h / 480.0, synth.CX * w / 640.0, synth.CY * h / 480.0
    host = []
    for f, d in zip(frames, depths):
        k, desc = orc.extract(f, **okw)
        host.append((k, desc, orc.unproject(k, d, fx=fx, fy=fy, cx=cx, cy=cy)[0]))
    assert min(len(x[0]) for x in host) > 1900
    ctx = ob.Context(width=w, height=h, nfeatures=2000, max_frames=n, fx=fx, fy=fy, cx=cx, cy=cy)
    try:
        assert ctx.K >= 2000
        cov = -1.0                              # latched once per context, like the reference's process-static (quirk Q7)
        for cross in (True, False):
            ctx.extract_batch(frames, depths)
            ctx.match_pairs(np.array([[0, 1], [1, 2], [2, 0]], np.int32), 0.8, cross_check=cross)
            ctx.ransac_pairs(3, seed=11)
            for p, (a, b) in enumerate([(0, 1), (1, 2), (2, 0)]):
                if p == 0 and cross:
                    for s in range(n):
                        k, d, xyz = ctx.download_frame(s)
                        assert k.tobytes() == host[s][0].tobytes() and np.array_equal(d, host[s][1]) and np.array_equal(xyz, host[s][2])
                i1, e1, i2, e2 = ctx.download_knn(p)
                r1 = orc.knn2(host[a][1], host[b][1])
                assert np.array_equal(i1, r1[0]) and np.array_equal(e1, r1[1]) and np.array_equal(i2, r1[2]) and np.array_equal(e2, r1[3])
                m = orc.knn_match(host[a][1], host[b][1], 0.8, cross)
                r = orc.ransac_iterate(host[a][2], host[b][2], m, seed=11 + p, depth_cov=cov)
                cov = r["depth_cov"]
                _assert_pair(ctx, p, m, r)
                assert r["ok"]
    finally:
        ctx.close()


def test_fr1_distortion_feeds_the_unprojection(ob, orc, texture):
    """Core/frame.cpp:148-164 with the reference's own calibration (k1 != 0): depth at the distorted keypoint, mvuRight / mvKeys3Dc
    from mvKeysUn = cv::undistortPoints(mvKeys) — and everything downstream (RANSAC inliers, T12) follows."""
    n = 4
    frames = np.stack([synth.make_frame(texture, 80 + i) for i in range(n)])
    depths = np.stack([synth.make_depth(80 + i) for i in range(n)])
    k1, k2, p1, p2, k3 = FR1_DIST
    host, ref = _oracle_chain(orc, frames, depths, 0.8, True, 5, dist=FR1_DIST)
    plain = orc.unproject(host[0][0], depths[0])[0]
    assert np.abs(plain - host[0][2]).max() > 1e-3, "the distortion must move the 3D points"
    ctx = ob.Context(max_frames=n, k1=k1, k2=k2, p1=p1, p2=p2, k3=k3)
    try:
        ctx.track_sequence(frames, depths, 0.8, cross_check=True, seed=5)
        for s in range(n):
            k, d, xyz = ctx.download_frame(s)
            assert k.tobytes() == host[s][0].tobytes() and np.array_equal(d, host[s][1])
            assert np.array_equal(xyz, host[s][2]), f"frame {s}: mvKeys3Dc"
            xy_un, ur = ctx.download_keys_un(s)
            xy = np.stack([k["x"], k["y"]], 1)
            assert np.array_equal(xy_un, orc.undistort_points(xy, synth.FX, synth.FY, synth.CX, synth.CY, FR1_DIST)), f"frame {s}: mvKeysUn"
            assert np.array_equal(ur, host[s][3]), f"frame {s}: mvuRight"
        for p, (m, r) in enumerate(ref):
            _assert_pair(ctx, p, m, r)
    finally:
        ctx.close()
    # k1 == 0: mvKeysUn = mvKeys (frame.cpp:288-291)
    ctx = ob.Context(max_frames=1)
    try:
        ctx.extract_batch(frames[:1], depths[:1])
        k, _, _ = ctx.download_frame(0)
        xy_un, _ = ctx.download_keys_un(0)
        assert np.array_equal(xy_un, np.stack([k["x"], k["y"]], 1))
    finally:
        ctx.close()


def test_config3_pipelined_512_frames_vs_oracle_on_72_pairs(ob, orc, texture):
    """The timed path (orbf_track_sequence over 512 frames: ramped chunks on 4 worker streams, zero-copy depth when pinned, lazy sample
    tables, hypothesis waves) against the oracle: every distinct (frame, frame) pair kind, 72 pairs byte for byte — the first 24, 24
    around chunk boundaries in the middle, the last 24 — and every frame's keypoints / descriptors / 3D points."""
    nb = 8
    base_f = [synth.make_frame(texture, i) for i in range(nb)]
    base_d = [synth.make_depth(i) for i in range(nb)]
    n = 512
    seq = list(range(nb)) + list(range(nb - 2, 0, -1))                   # ping-pong: 14 distinct consecutive pair kinds
    ids = [seq[i % len(seq)] for i in range(n)]
    frames = np.stack([base_f[j] for j in ids]); depths = np.stack([base_d[j] for j in ids])
    ext = [orc.extract(f) for f in base_f]
    xyz = [orc.unproject(ext[j][0], base_d[j])[0] for j in range(nb)]
    ctx = ob.Context(max_frames=n, max_pairs=n)
    try:
        import torch
        hg = torch.from_numpy(frames).pin_memory(); hd = torch.from_numpy(depths.view(np.int16)).pin_memory()
        ctx.track_sequence(hg.numpy(), hd.numpy().view(np.uint16), 0.8, cross_check=True, seed=42)
        fc = ctx.frame_counts(n)
        assert np.array_equal(fc, np.array([len(ext[j][0]) for j in ids]))
        for s in list(range(0, 20)) + list(range(250, 262)) + list(range(n - 10, n)):
            k, d, p3 = ctx.download_frame(s)
            assert k.tobytes() == ext[ids[s]][0].tobytes() and np.array_equal(d, ext[ids[s]][1]) and np.array_equal(p3, xyz[ids[s]]), f"frame {s}"
        matches = {}
        for a, b in set(zip(ids[:-1], ids[1:])):
            matches[(a, b)] = orc.knn_match(ext[a][1], ext[b][1], 0.8, True)
        assert len(matches) == 14
        mc = ctx.match_counts(n - 1)
        assert np.array_equal(mc, np.array([len(matches[(ids[p], ids[p + 1])]) for p in range(n - 1)])), "match counts of all 511 pairs"
        cov = orc.ransac_iterate(xyz[ids[0]], xyz[ids[1]], matches[(ids[0], ids[1])], seed=42)["depth_cov"]
        check = list(range(0, 24)) + list(range(244, 268)) + list(range(n - 1 - 24, n - 1))
        kinds = set()
        for p in check:
            a, b = ids[p], ids[p + 1]
            kinds.add((a, b))
            r = orc.ransac_iterate(xyz[a], xyz[b], matches[(a, b)], seed=42 + p, depth_cov=cov)
            _assert_pair(ctx, p, matches[(a, b)], r)
        assert len(kinds) == 14 and len(check) == 72
    finally:
        ctx.close()
