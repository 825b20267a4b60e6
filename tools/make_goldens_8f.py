#!/usr/bin/env python
"""Golden vectors for the §8f geometry entry points, minted from cv2 (the reference delegates exactly these computations to OpenCV):
cv2.undistortPoints with P = K (Core/frame.cpp:302) and cv::Mat products via cv2.gemm (Odometry/odometry.cpp:83, Features/matcher.cpp:229).
    python tools/make_goldens_8f.py      # writes tests/golden/geometry_cv2.npz"""
from pathlib import Path

import cv2
import numpy as np

OUT = Path(__file__).resolve().parent.parent / "tests" / "golden" / "geometry_cv2.npz"
FX, FY, CX, CY = 517.3, 516.5, 318.6, 255.3


def main():
    rng = np.random.default_rng(20260101)
    K = np.array([[FX, 0, CX], [0, FY, CY], [0, 0, 1]], np.float32)
    dists = np.array([[0.2624, -0.9531, -0.0054, 0.0026, 1.1633], [0.2312, -0.7849, -0.0033, -0.0001, 0.9172], [-0.3, 0.1, 0.001, -0.002, 0.0]], np.float32)
    pts = np.stack([rng.uniform(-5, 645, 600), rng.uniform(-5, 485, 600)], 1).astype(np.float32)
    pts[:4] = [[0, 0], [640, 0], [0, 480], [640, 480]]
    und = np.stack([cv2.undistortPoints(pts.reshape(-1, 1, 2), K, d, None, K).reshape(-1, 2) for d in dists])
    # trajectory: T12 chain composed by cv2.gemm exactly as `T12 * pF1->GetPose()`
    T = np.tile(np.eye(4, dtype=np.float32), (60, 1, 1))
    for k in range(60):
        a = rng.normal(0, 0.03, 3)
        Rx = np.array([[1, 0, 0], [0, np.cos(a[0]), -np.sin(a[0])], [0, np.sin(a[0]), np.cos(a[0])]])
        Ry = np.array([[np.cos(a[1]), 0, np.sin(a[1])], [0, 1, 0], [-np.sin(a[1]), 0, np.cos(a[1])]])
        Rz = np.array([[np.cos(a[2]), -np.sin(a[2]), 0], [np.sin(a[2]), np.cos(a[2]), 0], [0, 0, 1]])
        T[k, :3, :3] = (Rz @ Ry @ Rx).astype(np.float32); T[k, :3, 3] = rng.normal(0, 0.05, 3).astype(np.float32)
    pose0 = np.eye(4, dtype=np.float32); pose0[:3, 3] = [0.25, -0.5, 1.0]
    poses = [pose0]
    for k in range(60):
        poses.append(cv2.gemm(T[k], poses[-1], 1.0, None, 0.0))
    # Fuse projection: Rcw * p3Dw + tcw
    R = T[7, :3, :3].copy(); t = T[7, :3, 3].copy()
    pw = (rng.normal(0, 1, (300, 3)) * [1.0, 0.8, 0.5] + [0, 0, 2.5]).astype(np.float32)
    pc = np.stack([cv2.gemm(R, p.reshape(3, 1), 1.0, t.reshape(3, 1), 1.0)[:, 0] for p in pw])
    np.savez_compressed(OUT, opencv=np.array(cv2.__version__), intrinsics=np.array([FX, FY, CX, CY], np.float32), dists=dists, pts=pts, undistorted=und,
                        T12=T, pose0=pose0, poses=np.stack(poses), fuse_R=R, fuse_t=t, fuse_pw=pw, fuse_pc=pc)
    print(OUT, OUT.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
