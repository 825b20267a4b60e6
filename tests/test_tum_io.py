"""TUM association / trajectory files (reference Utils/utils.cpp:16-38, System/tracking.cpp:544-580)."""
import importlib.util
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
spec = importlib.util.spec_from_file_location("orbfront_tum", ROOT / "adaptive-rgbd-localization-mappig_b200" / "tum.py")
tum = importlib.util.module_from_spec(spec); spec.loader.exec_module(tum)


def test_associations_roundtrip(tmp_path):
    p = tmp_path / "assoc.txt"
    p.write_text("1305031102.175304 rgb/1305031102.175304.png 1305031102.160407 depth/1305031102.160407.png\n\n"
                 "1305031102.211214 rgb/1305031102.211214.png 1305031102.226738 depth/1305031102.226738.png\n")
    ts, rgb, dep = tum.load_associations(p)
    assert ts.tolist() == [1305031102.175304, 1305031102.211214]
    assert rgb == ["rgb/1305031102.175304.png", "rgb/1305031102.211214.png"] and dep[1] == "depth/1305031102.226738.png"


def _rot(rng):
    q = rng.normal(size=4); q /= np.linalg.norm(q)
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def test_trajectory_format_and_values(tmp_path):
    import cv2
    rng = np.random.default_rng(0)
    poses = []
    for _ in range(50):
        T = np.eye(4, dtype=np.float32); T[:3, :3] = _rot(rng).astype(np.float32); T[:3, 3] = rng.normal(size=3).astype(np.float32)
        poses.append(T)
    ts = 1305031102.0 + np.arange(50) * 0.0333
    p = tmp_path / "traj.txt"
    tum.save_trajectory(p, ts, poses)
    t2, xyz, q = tum.load_trajectory(p)
    assert np.allclose(t2, ts, atol=1e-6)
    for k, T in enumerate(poses):
        Rwc = np.ascontiguousarray(T[:3, :3].T)
        twc = cv2.gemm(Rwc, np.ascontiguousarray(T[:3, 3:4]), -1.0, None, 0.0)[:, 0]          # -Rwc * tcw as cv::Mat evaluates it
        assert np.array_equal(np.float32(xyz[k]), np.float32([float(f"{v:.9f}") for v in twc]))
        # the quaternion rotates like Rwc (sign-free check) and has unit norm
        x, y, z, w = q[k]
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                      [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
        assert np.allclose(R, Rwc, atol=2e-6) and abs(np.linalg.norm(q[k]) - 1) < 1e-6
    line = p.read_text().splitlines()[0].split()
    assert len(line) == 8 and len(line[0].split(".")[1]) == 6 and all(len(v.split(".")[1]) == 9 for v in line[1:])


def test_quaternion_branches():
    # trace <= 0 branches of Eigen's conversion: rotations by pi about each axis
    for axis in range(3):
        R = -np.eye(3); R[axis, axis] = 1
        q = tum.quaternion_from_rotation(R)
        want = np.zeros(4, np.float32); want[axis] = 1
        assert np.array_equal(np.abs(q), want)


def test_cpp_header_equals_the_python_host_logic(tmp_path):
    """include/orbfront_tum.hpp (LoadImages / SaveTrajectory for a C++ host: Utils/utils.cpp:16-38, System/tracking.cpp:566-577) against
    tum.py: the same association entries (blank lines, no trailing newline) and a byte-identical trajectory file over random poses, the
    trace <= 0 quaternion branches included; a file that cannot be opened throws instead of the reference's endless `while (!eof())`."""
    import struct
    import subprocess
    exe = tmp_path / "tum_io_demo"
    r = subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", "-Werror", f"-I{ROOT / 'include'}", "-o", str(exe), str(ROOT / "tests" / "cpp" / "tum_io_demo.cpp")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assoc = tmp_path / "assoc.txt"
    assoc.write_text("1305031102.175304 rgb/1305031102.175304.png 1305031102.160407 depth/1305031102.160407.png\n\n"
                     "1305031102.211214 rgb/1305031102.211214.png 1305031102.226738 depth/1305031102.226738.png\n"
                     "1305031102.243211 rgb/c.png 1305031102.262886 depth/c.png")
    rng = np.random.default_rng(1)
    poses = []
    for k in range(200):
        T = np.eye(4, dtype=np.float32); T[:3, :3] = _rot(rng).astype(np.float32); T[:3, 3] = (rng.normal(size=3) * 10.0 ** int(rng.integers(-3, 3))).astype(np.float32)
        poses.append(T)
    for axis in range(3):                                        # rotations by pi: the trace <= 0 branches
        T = np.eye(4, dtype=np.float32); T[:3, :3] = -np.eye(3); T[axis, axis] = 1; T[:3, 3] = [1, -2, 3]
        poses.append(T)
    ts = 1305031102.0 + np.arange(len(poses)) * 0.0333333
    (tmp_path / "poses.bin").write_bytes(struct.pack("i", len(poses)) + ts.astype(np.float64).tobytes() + np.stack(poses).astype(np.float32).tobytes())
    r = subprocess.run([str(exe), str(assoc), str(tmp_path / "poses.bin"), str(tmp_path / "traj_cpp.txt"), str(tmp_path / "assoc_cpp.txt")],
                       capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stderr)
    tum.save_trajectory(tmp_path / "traj_py.txt", ts, poses)
    assert (tmp_path / "traj_cpp.txt").read_bytes() == (tmp_path / "traj_py.txt").read_bytes()
    t_py, rgb_py, dep_py = tum.load_associations(assoc)
    rows = [l.split("|") for l in (tmp_path / "assoc_cpp.txt").read_text().splitlines()]
    assert [float(a[0]) for a in rows] == t_py.tolist() and [a[1] for a in rows] == rgb_py and [a[2] for a in rows] == dep_py and len(rows) == 3


def test_load_images_equals_the_reference_source(tmp_path):
    """orbf::LoadImages (include/orbfront_tum.hpp) and tum.load_associations against the reference's OWN LoadImages — Utils/utils.cpp
    compiled verbatim into oracle/_ref/ref_utils_demo (oracle/Makefile): the same timestamps (to the last bit), file names and entry
    count on well-formed files with blank lines, CRLF line ends, a missing last newline, and on lines with missing trailing fields (empty
    names, like the reference).  Not compared: whitespace-only lines, where the reference pushes an uninitialised double."""
    import subprocess
    sys.path.insert(0, str(ROOT))
    from oracle import ref
    if not ref.available() or not ref.UTILS_DEMO.exists():
        import pytest
        pytest.skip("oracle/_ref/ref_utils_demo is not built and the reference checkout is absent")
    exe = tmp_path / "tum_io_demo"
    r = subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", "-Werror", f"-I{ROOT / 'include'}", "-o", str(exe), str(ROOT / "tests" / "cpp" / "tum_io_demo.cpp")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    (tmp_path / "poses.bin").write_bytes(b"\0\0\0\0")
    rng = np.random.default_rng(3)
    well = "".join(f"{1305031102.0 + 0.0333 * k + rng.random() * 1e-3:.6f} rgb/{k}.png {1305031102.0 + 0.0333 * k:.6f} depth/{k}.png\n" + ("\n" if k % 7 == 0 else "")
                   for k in range(100))
    cases = {"well": well, "no_last_newline": well.rstrip("\n"), "crlf": well.replace("\n\n", "\n").replace("\n", "\r\n"),
             "missing_fields": "1.5 rgb/a.png 1.6 depth/a.png\n2.5 rgb/b.png 2.6\n3.5 rgb/c.png\n4.5\n5.5 rgb/e.png 5.6 depth/e.png\n", "empty": ""}
    for name, text in cases.items():
        p = tmp_path / f"{name}.txt"
        p.write_bytes(text.encode())
        r = subprocess.run([str(ref.UTILS_DEMO), str(p), str(tmp_path / f"{name}.ref")], capture_output=True, text=True, timeout=60)
        assert r.returncode == 0, (name, r.stderr)
        r = subprocess.run([str(exe), str(p), str(tmp_path / "poses.bin"), str(tmp_path / "t.txt"), str(tmp_path / f"{name}.cpp")], capture_output=True, text=True)
        assert r.returncode == 0, (name, r.returncode, r.stderr)
        want = (tmp_path / f"{name}.ref").read_text()
        assert (tmp_path / f"{name}.cpp").read_text() == want, name
        if name in ("well", "no_last_newline", "crlf"):
            rows = [l.split("|") for l in want.splitlines()]
            ts, rgb, dep = tum.load_associations(p)
            assert len(rows) == 100 and ts.tolist() == [float(a[0]) for a in rows] and rgb == [a[1] for a in rows] and dep == [a[2] for a in rows], name
    assert (tmp_path / "empty.ref").read_text() == ""
