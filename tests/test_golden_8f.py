"""Committed cv2 golden vectors for the §8f geometry entry points (tools/make_goldens_8f.py): cv2.undistortPoints with P = K and
cv::Mat products via cv2.gemm.  CPU: the oracle reproduces them bit for bit (so parity does not depend on cv2 being importable
where the tests run); GPU: so does the CUDA path."""
from pathlib import Path

import numpy as np
import pytest

G = np.load(Path(__file__).resolve().parent / "golden" / "geometry_cv2.npz")
FX, FY, CX, CY = [float(v) for v in G["intrinsics"]]


def test_oracle_undistort_equals_golden(orc):
    for k in range(len(G["dists"])):
        assert np.array_equal(orc.undistort_points(G["pts"], FX, FY, CX, CY, G["dists"][k]), G["undistorted"][k])


def test_oracle_trajectory_equals_golden(orc):
    assert np.array_equal(orc.compose_trajectory(G["T12"], G["pose0"]), G["poses"])


def test_oracle_fuse_projection_equals_golden(orc):
    """Rcw * p3Dw + tcw as cv::gemm evaluates it: checked through orc_fuse_search on a one-feature frame placed at the projection."""
    R, t, pw, pc = G["fuse_R"], G["fuse_t"], G["fuse_pw"], G["fuse_pc"]
    cam = np.array([FX, FY, CX, CY, 40.0, -1e9, 1e9, -1e9, 1e9], np.float32)
    f32 = np.float32
    for i in range(len(pw)):
        if pc[i, 2] <= 0.05:
            continue
        invz = f32(1) / pc[i, 2]
        u = f32(f32(FX) * f32(pc[i, 0] * invz)) + f32(CX); v = f32(f32(FY) * f32(pc[i, 1] * invz)) + f32(CY)
        d = np.zeros((1, 32), np.uint8)
        # a feature exactly at the golden projection is found (distance 0, zero reprojection error); one 4 px away is not
        best, _ = orc.fuse_search(R, t, cam, [u], [v], [-1.0], d, pw[i:i + 1], d, [1], radius=3.0)
        assert best[0] == 0
        best, _ = orc.fuse_search(R, t, cam, [u + f32(4)], [v], [-1.0], d, pw[i:i + 1], d, [1], radius=3.0)
        assert best[0] == -1


@pytest.mark.gpu
def test_cuda_equals_golden(ob):
    ctx = ob.Context(max_frames=2)
    for k in range(len(G["dists"])):
        assert np.array_equal(ctx.undistort_points(G["pts"], FX, FY, CX, CY, G["dists"][k]), G["undistorted"][k])
