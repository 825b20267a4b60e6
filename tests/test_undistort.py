"""Frame::UndistortKeyPoints / ComputeImageBounds (reference Core/frame.cpp:286-343): cv::undistortPoints with
P = K.  The oracle restates OpenCV's cvUndistortPointsInternal and is pinned bit-exactly against cv2.undistortPoints; the CUDA kernel
is compared bit-exactly with the oracle."""
import numpy as np
import pytest

FX, FY, CX, CY = 517.3, 516.5, 318.6, 255.3                                  # TUM FR1 (common.h:35-38)
DISTS = [np.array([0.2624, -0.9531, -0.0054, 0.0026, 1.1633], np.float32),   # TUM FR1 distortion
         np.array([0.2312, -0.7849, -0.0033, -0.0001, 0.9172], np.float32),  # TUM FR2
         np.array([-0.3, 0.1, 0.001, -0.002, 0.0], np.float32)]


def _points(seed, n=4000):
    rng = np.random.default_rng(seed)
    pts = np.stack([rng.uniform(-5, 645, n), rng.uniform(-5, 485, n)], 1).astype(np.float32)
    pts[:4] = [[0, 0], [640, 0], [0, 480], [640, 480]]                         # the ComputeImageBounds corners
    return pts


@pytest.mark.parametrize("k", range(len(DISTS)))
def test_oracle_is_cv2_undistort_points(orc, k):
    import cv2
    pts = _points(k)
    K = np.array([[FX, 0, CX], [0, FY, CY], [0, 0, 1]], np.float32)
    ref = cv2.undistortPoints(pts.reshape(-1, 1, 2), K, DISTS[k], None, K).reshape(-1, 2)
    got = orc.undistort_points(pts, FX, FY, CX, CY, DISTS[k])
    assert np.array_equal(got, ref)
    assert len(orc.undistort_points(pts[:0], FX, FY, CX, CY, DISTS[k])) == 0


@pytest.mark.gpu
@pytest.mark.parametrize("k", range(len(DISTS)))
def test_cuda_matches_oracle(ob, orc, k):
    pts = _points(10 + k, 20000)
    ctx = ob.Context(max_frames=2)
    assert np.array_equal(ctx.undistort_points(pts, FX, FY, CX, CY, DISTS[k]), orc.undistort_points(pts, FX, FY, CX, CY, DISTS[k]))
    assert len(ctx.undistort_points(pts[:0], FX, FY, CX, CY, DISTS[k])) == 0


def test_oracle_unprojects_the_undistorted_keypoint(orc):
    """Core/frame.cpp:148-164: depth at the truncated DISTORTED keypoint, mvuRight / mvKeys3Dc from mvKeysUn — restated here with numpy
    float32 operations on top of cv2.undistortPoints."""
    import cv2
    rng = np.random.default_rng(3)
    n = 500
    kps = np.zeros(n, orc.KEYPOINT_DT)
    kps["x"] = rng.uniform(19, 620, n).astype(np.float32); kps["y"] = rng.uniform(19, 460, n).astype(np.float32)
    depth = rng.integers(0, 20000, size=(480, 640)).astype(np.uint16)
    depth[rng.random((480, 640)) < 0.1] = 0
    K = np.array([[FX, 0, CX], [0, FY, CY], [0, 0, 1]], np.float32)
    xy = np.stack([kps["x"], kps["y"]], 1)
    un = cv2.undistortPoints(xy.reshape(-1, 1, 2), K, DISTS[0], None, K).reshape(-1, 2)
    f32 = np.float32
    z = depth[kps["y"].astype(np.int64), kps["x"].astype(np.int64)].astype(f32) * f32(1.0 / 5000.0)
    ok = z > 0
    invfx, invfy = f32(1.0) / f32(FX), f32(1.0) / f32(FY)
    with np.errstate(divide="ignore", invalid="ignore"):
        X = np.where(ok, (un[:, 0] - f32(CX)) * z * invfx, f32(0)); Y = np.where(ok, (un[:, 1] - f32(CY)) * z * invfy, f32(0))
        ur = np.where(ok, un[:, 0] - f32(40.0) / z, f32(-1))
    xyz, got_ur = orc.unproject(kps, depth, dist=DISTS[0])
    assert np.array_equal(xyz, np.stack([X, Y, np.where(ok, z, f32(0))], 1).astype(f32)) and np.array_equal(got_ur, ur.astype(f32))
    plain, _ = orc.unproject(kps, depth)                      # k1 == 0 shortcut
    assert np.array_equal(plain, orc.unproject(kps, depth, dist=np.zeros(5, np.float32))[0]) and not np.array_equal(plain, xyz)
