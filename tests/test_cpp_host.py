"""The C++ host mirror of the reference interface (include/orbfront_host.hpp: ORBextractor, Extractor, Frame, Matcher,
Ransac, Kabsch, Odometry) driven like the reference's tracking loop by tests/cpp/host_api_demo.cpp.

  not gpu : the headers compile with g++ -Wall against include/ only, link against liborbfront_b200.so, and the program
            refuses to run without a CUDA device (exit 3: no CPU fallback);
  gpu     : its outputs equal the oracle chained the same way (System/tracking.cpp:38-46,193-208)."""
import struct
import subprocess
from pathlib import Path

import numpy as np
import pytest

import synth

ROOT = Path(__file__).resolve().parent.parent
PKG = ROOT / "adaptive-rgbd-localization-mappig_b200"


@pytest.fixture(scope="module")
def demo(ob, tmp_path_factory):
    ob.lib()                                                     # builds liborbfront_b200.so if needed
    exe = tmp_path_factory.mktemp("cpp") / "host_api_demo"
    cmd = ["g++", "-std=c++17", "-O2", "-Wall", "-Werror", f"-I{ROOT / 'include'}", "-o", str(exe), str(ROOT / "tests" / "cpp" / "host_api_demo.cpp"),
           f"-L{PKG}", "-lorbfront_b200", f"-Wl,-rpath,{PKG}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def write_input(path, frames, depths):
    n, h, w = frames.shape
    with open(path, "wb") as f:
        f.write(struct.pack("3i", n, w, h)); f.write(frames.tobytes()); f.write(depths.tobytes())


def test_host_mirror_compiles_and_refuses_without_gpu(demo, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    write_input(tmp_path / "in.raw", np.zeros((1, 480, 640), np.uint8), np.zeros((1, 480, 640), np.uint16))
    r = subprocess.run([str(demo), str(tmp_path / "in.raw"), str(tmp_path / "out.raw")], capture_output=True, text=True)
    assert r.returncode == 3 and "CUDA" in r.stderr, (r.returncode, r.stderr)


@pytest.mark.gpu
def test_host_mirror_matches_oracle(demo, ob, orc, texture, tmp_path):
    n = 4
    frames = np.stack([synth.make_frame(texture, 50 + i) for i in range(n)])
    depths = np.stack([synth.make_depth(50 + i) for i in range(n)])
    write_input(tmp_path / "in.raw", frames, depths)
    r = subprocess.run([str(demo), str(tmp_path / "in.raw"), str(tmp_path / "out.raw")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    buf = (tmp_path / "out.raw").read_bytes()
    pos = 0

    def take(dtype, count):
        nonlocal pos
        a = np.frombuffer(buf, dtype, count, pos); pos += a.nbytes
        return a

    prev = None; cov = -1.0; T12s = []; hostd = []
    for i in range(n):
        N = int(take(np.int32, 1)[0])
        k = take(ob.KEYPOINT_DT, N); d = take(np.uint8, N * 32).reshape(N, 32); xyz = take(np.float32, N * 3).reshape(N, 3)
        ko, do = orc.extract(frames[i]); xo, _ = orc.unproject(ko, depths[i])
        assert k.tobytes() == ko.tobytes() and np.array_equal(d, do) and np.array_equal(xyz, xo), f"frame {i}"
        if prev is not None:
            nm = int(take(np.int32, 1)[0]); m = take(ob.DMATCH_DT, nm)
            ok = int(take(np.int32, 1)[0]); rmse = float(take(np.float32, 1)[0]); T = take(np.float32, 16).reshape(4, 4)
            ni = int(take(np.int32, 1)[0]); inl = take(ob.DMATCH_DT, ni)
            mo = orc.knn_match(prev[0], do, 0.8, True)
            assert m.tobytes() == mo.tobytes(), f"pair {i - 1}: matches"
            ro = orc.ransac_iterate(prev[1], xo, mo, seed=42 + (i - 1), depth_cov=cov)     # Ransac::Seed() advances per call
            cov = ro["depth_cov"]
            assert ok == int(ro["ok"]) and inl.tobytes() == ro["inliers"].tobytes() and rmse == ro["rmse"]
            assert np.abs(T - ro["T12"]).max() <= 1e-5
            T12s.append(T)
        prev = (do, xo); hostd.append((do, xo))
    dist = int(take(np.int32, 1)[0])
    d0 = orc.extract(frames[0])[1]
    assert dist == int(np.unpackbits(d0[0] ^ d0[1]).sum())                                   # Matcher::DescriptorDistance
    T = take(np.float32, 16).reshape(4, 4)
    assert np.abs(T[:3, :3] - np.array([[0, -1, 0], [1, 0, 0], [0, 0, 1]], np.float32)).max() < 1e-4
    assert np.abs(T[:3, 3] - np.array([0.1, -0.2, 0.3], np.float32)).max() < 1e-4
    T2 = take(np.float32, 16).reshape(4, 4)
    assert np.array_equal(T2, T)                                                             # Kabsch::Compute(MatrixXf, MatrixXf) = the same call
    poses = take(np.float32, 16 * n).reshape(n, 4, 4)                                        # Odometry::Compute: T12 * pF1->GetPose() (odometry.cpp:82-86)
    assert np.array_equal(poses, orc.compose_trajectory(np.stack(T12s)))
    nl, pw, ph = (int(x) for x in take(np.int32, 3))                                         # ORBextractor::mvImagePyramid[2] of the last frame
    lvl = take(np.uint8, pw * ph).reshape(ph, pw)
    assert nl == 8 and np.array_equal(lvl, orc.pyramid(frames[n - 1])[2])
    # Matcher::KnnMatch(KeyFrame*, Frame&, .): landmark j + 1 at every even feature j of the keyframe, (j % 10 == 0) bad, feature 7 of F2 taken
    nk = int(take(np.int32, 1)[0]); mk = take(ob.DMATCH_DT, nk); n_out = int(take(np.int32, 1)[0])
    lm1 = np.array([j + 1 if j % 2 == 0 else 0 for j in range(len(hostd[0][0]))]); lm2 = np.zeros(len(hostd[1][0]), np.int64); lm2[7] = 0x7fff
    want = orc.knn_match_keyframe(hostd[0][0], hostd[1][0], 0.8, lm1, lambda p: (p - 1) % 10 == 0, lm2)
    assert mk.tobytes() == want.tobytes() and 0 < nk < len(orc.knn_match(hostd[0][0], hostd[1][0], 0.8, False)) and n_out == nk
    # Ransac(KeyFrame*, KeyFrame*, matches).Iterate() and its clouds
    ok2 = int(take(np.int32, 1)[0]); Tb = take(np.float32, 16).reshape(4, 4); nib = int(take(np.int32, 1)[0]); inlb = take(ob.DMATCH_DT, nib)
    nc = int(take(np.int32, 1)[0]); cs = take(np.float32, nc * 4).reshape(nc, 4); ct = take(np.float32, nc * 4).reshape(nc, 4)
    m01 = orc.knn_match(hostd[0][0], hostd[1][0], 0.8, True)
    rb = orc.ransac_iterate(hostd[0][1], hostd[1][1], m01, seed=100, depth_cov=cov)
    assert ok2 == int(rb["ok"]) and inlb.tobytes() == rb["inliers"].tobytes() and np.abs(Tb - rb["T12"]).max() <= 1e-5
    ws, wt = orc.ransac_clouds(hostd[0][1], hostd[1][1], m01)
    assert nc == len(ws) > 20 and np.array_equal(cs, ws) and np.array_equal(ct, wt)
    assert tuple(int(x) for x in take(np.int32, 3)) == (0, 0, 0)                             # fewer than minInlierTh matches: early return
    th = np.full(9, 20.0)
    for i in range(n):                                                                       # Extractor(FAST, ., ADAPTIVE)
        N = int(take(np.int32, 1)[0]); k = take(ob.KEYPOINT_DT, N)
        assert k.tobytes() == orc.adaptive_detect(frames[i], th, retain_best=1000)[0].tobytes(), f"adaptive frame {i}"
    Na = int(take(np.int32, 1)[0]); ka = take(ob.KEYPOINT_DT, Na); xa = take(np.float32, Na * 3).reshape(Na, 3); ua = take(np.float32, Na)
    tha = np.full(9, 20.0)                                                                   # Frame::ExtractFeatures on the adaptive route
    kref = orc.adaptive_detect(frames[0], tha, retain_best=1000)[0]
    xref, uref = orc.unproject(kref, depths[0])
    assert ka.tobytes() == kref.tobytes() and np.array_equal(xa, xref) and np.array_equal(ua, uref)
    assert np.array_equal(take(np.float64, 9), th)
    assert pos == len(buf)


@pytest.fixture(scope="module")
def ocv_demo(ob, tmp_path_factory):
    """The ORBF_WITH_OPENCV overload set (cv::InputArray / cv::OutputArray / cv::Mat signatures of the reference) compiled against the
    stand-in header tests/cpp/stub/opencv2/core.hpp — this image has no OpenCV headers."""
    ob.lib()
    exe = tmp_path_factory.mktemp("cpp_ocv") / "opencv_signatures"
    cmd = ["g++", "-std=c++17", "-O2", "-Wall", "-Werror", f"-I{ROOT / 'include'}", f"-I{ROOT / 'tests' / 'cpp' / 'stub'}", "-o", str(exe),
           str(ROOT / "tests" / "cpp" / "opencv_signatures.cpp"), f"-L{PKG}", "-lorbfront_b200", f"-Wl,-rpath,{PKG}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_opencv_overloads_compile_and_refuse_without_gpu(ocv_demo, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    write_input(tmp_path / "in.raw", np.zeros((1, 480, 640), np.uint8), np.zeros((1, 480, 640), np.uint16))
    r = subprocess.run([str(ocv_demo), str(tmp_path / "in.raw"), str(tmp_path / "out.raw")], capture_output=True, text=True)
    assert r.returncode == 3 and "CUDA" in r.stderr, (r.returncode, r.stderr)


@pytest.mark.gpu
def test_opencv_overloads_match_oracle(ocv_demo, ob, orc, texture, tmp_path):
    frame = synth.make_frame(texture, 77)
    write_input(tmp_path / "in.raw", frame[None], np.zeros((1, 480, 640), np.uint16))
    r = subprocess.run([str(ocv_demo), str(tmp_path / "in.raw"), str(tmp_path / "out.raw")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    buf = (tmp_path / "out.raw").read_bytes()
    n = int(np.frombuffer(buf, np.int32, 1, 0)[0])
    k = np.frombuffer(buf, ob.KEYPOINT_DT, n, 4); d = np.frombuffer(buf, np.uint8, n * 32, 4 + 28 * n).reshape(n, 32)
    same, dist, kept = (int(x) for x in np.frombuffer(buf, np.int32, 3, 4 + 60 * n))
    ko, do = orc.extract(frame)
    assert k.tobytes() == ko.tobytes() and np.array_equal(d, do)
    assert same == 1 and kept == 3 and dist == int(np.unpackbits(do[0] ^ do[1]).sum())


@pytest.fixture(scope="module")
def eigen_demo(ob, tmp_path_factory):
    """Kabsch::Compute(const Eigen::MatrixXf&, const Eigen::MatrixXf&) -> Eigen::Matrix4f (ORBF_WITH_EIGEN) compiled against the stand-in
    header tests/cpp/stub/Eigen/Core — this image has no Eigen headers."""
    ob.lib()
    exe = tmp_path_factory.mktemp("cpp_eigen") / "eigen_signatures"
    cmd = ["g++", "-std=c++17", "-O2", "-Wall", "-Werror", f"-I{ROOT / 'include'}", f"-I{ROOT / 'tests' / 'cpp' / 'stub'}", "-o", str(exe),
           str(ROOT / "tests" / "cpp" / "eigen_signatures.cpp"), f"-L{PKG}", "-lorbfront_b200", f"-Wl,-rpath,{PKG}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_eigen_overload_compiles_and_refuses_without_gpu(eigen_demo):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = subprocess.run([str(eigen_demo)], capture_output=True, text=True)
    assert r.returncode == 3 and "CUDA" in r.stderr, (r.returncode, r.stderr)


@pytest.mark.gpu
def test_eigen_overload_matches_oracle(eigen_demo, orc):
    r = subprocess.run([str(eigen_demo)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    T = np.array([float(x) for x in r.stdout.split()], np.float32).reshape(4, 4)
    i = np.arange(40)
    A = np.stack([0.1 * i, 0.05 * (i * i % 17), 1.0 + 0.02 * i], 1).astype(np.float32)
    B = np.stack([-A[:, 1] + np.float32(0.1), A[:, 0] - np.float32(0.2), A[:, 2] + np.float32(0.3)], 1).astype(np.float32)
    assert np.abs(T - orc.kabsch(A, B)).max() <= 1e-5
    assert np.abs(T[:3, :3] - np.array([[0, -1, 0], [1, 0, 0], [0, 0, 1]], np.float32)).max() < 1e-4
