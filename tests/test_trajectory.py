"""Odometry::Compute, RANSAC strategy (reference Odometry/odometry.cpp:78-90): composition rule
pose[k+1] = T12[k] * pose[k] and the inlier flags, along a device-resident sequence.  Bit-exact floats: the product is cv::Mat's
(cv::gemm small-matrix path), pinned against cv2.gemm."""
import numpy as np
import pytest

import synth


def test_oracle_composition_is_cv_gemm(orc):
    import cv2
    rng = np.random.default_rng(3)
    T = rng.normal(size=(40, 4, 4)).astype(np.float32)
    T[:, 3] = [0, 0, 0, 1]
    p0 = rng.normal(size=(4, 4)).astype(np.float32)
    got = orc.compose_trajectory(T, p0)
    ref = [p0]
    for k in range(len(T)):
        ref.append(cv2.gemm(T[k], ref[-1], 1.0, None, 0.0))           # T12 * pF1->GetPose()
    assert np.array_equal(got, np.stack(ref))
    assert np.array_equal(orc.compose_trajectory(T[:0]), np.eye(4, dtype=np.float32)[None])


@pytest.mark.gpu
def test_cuda_trajectory_and_flags(ob, orc, texture):
    n = 7
    frames = np.stack([synth.make_frame(texture, i) for i in range(n)])
    depths = np.stack([synth.make_depth(i) for i in range(n)])
    ctx = ob.Context(max_frames=n)
    ctx.track_sequence(frames, depths, 0.8, True, seed=42)
    summ = ctx.download_ransac_summary(n - 1)
    p0 = np.eye(4, dtype=np.float32); p0[:3, 3] = [0.1, -0.2, 0.3]
    poses, flags = ctx.compose_trajectory(n - 1, p0)
    ref = orc.compose_trajectory(np.stack([s for s in summ["T12"]]), p0)
    assert np.array_equal(poses, ref)
    assert not np.array_equal(poses[-1], p0)                          # the sequence does move
    assert flags.shape == (n, ctx.K) and flags[0].all()
    for p in range(n - 1):
        r = ctx.download_ransac(p)
        want = np.ones(ctx.K, np.uint8); want[r["inliers"]["trainIdx"]] = 0
        assert np.array_equal(flags[p + 1], want)
    # identity start, no flags
    poses2, none = ctx.compose_trajectory(n - 1, None, with_flags=False)
    assert none is None and np.array_equal(poses2, orc.compose_trajectory(np.stack([s for s in summ["T12"]])))
    poses3, _ = ctx.compose_trajectory(0, p0)
    assert np.array_equal(poses3, p0[None])
