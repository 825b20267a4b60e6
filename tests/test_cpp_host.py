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

    prev = None; cov = -1.0
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
        prev = (do, xo)
    dist = int(take(np.int32, 1)[0])
    d0 = orc.extract(frames[0])[1]
    assert dist == int(np.unpackbits(d0[0] ^ d0[1]).sum())                                   # Matcher::DescriptorDistance
    T = take(np.float32, 16).reshape(4, 4)
    assert np.abs(T[:3, :3] - np.array([[0, -1, 0], [1, 0, 0], [0, 0, 1]], np.float32)).max() < 1e-4
    assert np.abs(T[:3, 3] - np.array([0.1, -0.2, 0.3], np.float32)).max() < 1e-4
    th = np.full(9, 20.0)
    for i in range(n):                                                                       # Extractor(FAST, ., ADAPTIVE)
        N = int(take(np.int32, 1)[0]); k = take(ob.KEYPOINT_DT, N)
        assert k.tobytes() == orc.adaptive_detect(frames[i], th, retain_best=1000)[0].tobytes(), f"adaptive frame {i}"
    assert np.array_equal(take(np.float64, 9), th)
    assert pos == len(buf)
