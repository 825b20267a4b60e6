"""CPU tests of the RANSAC / Kabsch oracle and of the host+device restatements in csrc/replay.h
(libstdc++ std::sort and glibc rand()), pinned against the real libraries the reference links."""
import numpy as np
import pytest

import synth


def test_svd3_against_numpy(orc):
    rng = np.random.default_rng(0)
    for i in range(1500):
        A = rng.normal(size=(3, 3)).astype(np.float32)
        if i % 5 == 0:
            A[:, 2] = A[:, 0] * 2
        if i % 7 == 0:
            A *= 1e-6
        if i % 11 == 0:
            A = np.diag(rng.normal(size=3)).astype(np.float32)
        U, S, V = orc.svd3(A)
        scale = max(np.abs(A).max(), 1e-30)
        assert np.abs(U @ np.diag(S) @ V.T - A).max() / scale < 1e-5
        assert np.abs(U.T @ U - np.eye(3)).max() < 1e-5 and np.abs(V.T @ V - np.eye(3)).max() < 1e-5
        sn = np.linalg.svd(A.astype(np.float64), compute_uv=False)
        assert np.abs(S - sn).max() / max(sn.max(), 1e-30) < 1e-5
        assert S[0] >= S[1] >= S[2] >= 0
    U, S, V = orc.svd3(np.zeros((3, 3)))
    assert np.array_equal(U, np.eye(3)) and np.array_equal(V, np.eye(3)) and (S == 0).all()


def test_weighted_transform_recovers_rigid_motion(orc):
    src, dst, m, R, t = synth.rigid_pairs(seed=2, outlier_frac=0.0)
    P = src[m["queryIdx"]]; P = P[P[:, 2] > 0]; Q = (P.astype(np.float64) @ R.T + t).astype(np.float32)
    T = orc.weighted_transform(P, Q)
    assert np.abs(T[:3, :3] - R).max() < 1e-5 and np.abs(T[:3, 3] - t).max() < 2e-5
    assert np.array_equal(T[3], [0, 0, 0, 1])


def test_mahalanobis_against_numpy(orc):
    rng = np.random.default_rng(1)
    src, dst, m, R, t = synth.rigid_pairs(seed=3, outlier_frac=0.0)
    T = np.eye(4, dtype=np.float32); T[:3, :3] = R; T[:3, 3] = t
    cx = (3 * np.tan(58.0 / 180.0 * np.pi / 640)) ** 2; cy = (3 * np.tan(45.0 / 180.0 * np.pi / 480)) ** 2; cz = 1.3e-3
    checked = 0
    for k in range(200):
        p = src[m["queryIdx"][k]]; q = dst[m["trainIdx"][k]] + rng.normal(0, 0.004, 3).astype(np.float32)
        if not (p[2] > 0 and q[2] > 0):
            continue
        Td = T.astype(np.float64)
        mu = Td[:3, :3] @ p.astype(np.float64) + Td[:3, 3]
        dl = mu - q.astype(np.float64)
        got = orc.mahalanobis2(p, q, T, cz)
        if dl @ dl > 2 * (max(cx, cz) * 2):
            assert got == np.finfo(np.float64).max
            continue
        S = Td[:3, :3].T @ np.diag([cx * p[2], cy * p[2], cz]) @ Td[:3, :3] + np.diag([cx * q[2], cy * q[2], cz])
        ref = dl @ np.linalg.solve(S, dl)
        assert abs(got - ref) <= 1e-9 * max(1.0, abs(ref))
        checked += 1
    assert checked > 50
    assert orc.mahalanobis2(np.array([0, 0, np.nan], np.float32), dst[0], T, cz) == np.finfo(np.float64).max


def test_ransac_recovers_motion_and_respects_rules(orc):
    src, dst, m, R, t = synth.rigid_pairs(seed=4, outlier_frac=0.3)
    r = orc.ransac_iterate(src, dst, m, seed=42)
    assert r["ok"] and len(r["inliers"]) >= 0.6 * r["n_good"]
    assert np.abs(r["T12"][:3, :3] - R).max() < 5e-3 and np.abs(r["T12"][:3, 3] - t).max() < 2e-2
    assert r["n_good"] == len(m) - 10                        # 5 + 3 holes (z == 0) and 2 NaN targets filtered out
    d = r["good_sorted"]["distance"]
    assert (np.diff(d) >= 0).all()
    assert r["real_iters"] <= 200 and r["valid_iters"] >= 1
    inl = set(zip(r["inliers"]["queryIdx"].tolist(), r["inliers"]["trainIdx"].tolist()))
    assert inl <= set(zip(m["queryIdx"].tolist(), m["trainIdx"].tolist()))
    few = orc.ransac_iterate(src, dst, m[:19])
    assert not few["ok"] and few["rmse"] == 1e6 and np.array_equal(few["T12"], np.eye(4))


def test_ransac_early_exit_skips_iterations(orc):
    """> 80 % inliers on an accepted hypothesis ends the loop (ransac.cpp:242-247): few real iterations."""
    src, dst, m, _, _ = synth.rigid_pairs(seed=5, outlier_frac=0.02)
    r = orc.ransac_iterate(src, dst, m, seed=1)
    assert r["ok"] and r["real_iters"] < 20


def test_kabsch_oracle(orc):
    src, dst, m, R, t = synth.rigid_pairs(seed=6, outlier_frac=0.0)
    A = src[m["queryIdx"]]; A = A[A[:, 2] > 0][:280]; B = (A.astype(np.float64) @ R.T + t).astype(np.float32)
    T = orc.kabsch(A, B)
    assert np.abs(T[:3, :3] - R).max() < 1e-5 and np.abs(T[:3, 3] - t).max() < 1e-5
    assert np.array_equal(orc.kabsch(np.zeros((0, 3)), np.zeros((0, 3))), np.eye(4, dtype=np.float32))


# ---- csrc/replay.h pinned against libc / libstdc++ (host side of the shared library; no GPU needed) ----
def test_glibc_rand_replay(ob, orc):
    for seed in (0, 1, 42, 123456789, 2 ** 31 + 5, 2 ** 32 - 1):
        assert np.array_equal(ob.selftest_glibc_rand(seed, 3000), orc.libc_rand_sequence(seed, 3000)), seed


def test_sample_table_replay(ob, orc):
    for seed, M in ((42, 640), (7, 21), (9, 4), (3, 3), (5, 1000), (11, 5)):
        assert np.array_equal(ob.selftest_sample_table(seed, M), orc.sample_table_libc(seed, M)), (seed, M)
    t = orc.sample_table_libc(1, 50, 500, 4)
    assert (np.diff(t, axis=1) > 0).all() and t.min() >= 0 and t.max() < 50      # ascending unique ids
    assert np.median(t) < 25                                                        # min(r1, r2): biased to low ids


def _killer(n):
    """median-of-3 adversary: drives introsort to its depth limit so the heapsort fallback runs."""
    a = np.zeros(n, np.float32)
    k = n // 2
    for i in range(k):
        a[i] = i + 1 if i % 2 == 0 else k + i + (1 if k % 2 else 0)
        a[k + i] = 2 * (i + 1)
    return a


def test_std_sort_replay(ob, orc):
    rng = np.random.default_rng(0)
    for t in range(300):
        n = int(rng.integers(0, 1500))
        d = np.zeros(n, ob.DMATCH_DT); d["queryIdx"] = np.arange(n); d["trainIdx"] = rng.permutation(n) if n else 0
        mode = t % 5
        if mode == 0: d["distance"] = rng.integers(0, 64, n)
        elif mode == 1: d["distance"] = rng.integers(0, 3, n)
        elif mode == 2: d["distance"] = np.sort(rng.integers(0, 256, n))[::-1]
        elif mode == 3: d["distance"] = rng.random(n)
        else: d["distance"] = _killer(n) if n >= 2 else 0
        assert ob.selftest_introsort(d).tobytes() == orc.std_sort_dmatch(d).tobytes(), (t, n, mode)
    for n in (4096, 20000):            # deep recursion / heapsort path on adversarial input
        d = np.zeros(n, ob.DMATCH_DT); d["queryIdx"] = np.arange(n); d["distance"] = _killer(n)
        assert ob.selftest_introsort(d).tobytes() == orc.std_sort_dmatch(d).tobytes()


def test_warp_partition_rule_equals_libstdcxx(tmp_path):
    """The rule csrc/ransac.cu: warp_unguarded_partition is built on (k-th left stop exchanged with k-th right stop until the positions
    cross; cut from two prefix sums), run as 32 lockstep lanes by tests/cpp/warp_partition_model.cpp, against the REAL
    std::__unguarded_partition of this libstdc++ and csrc/replay.h's sequential restatement: cut and array contents identical over 20 000
    ranges with heavy ties, all-equal, sorted and reversed content (Odometry/ransac.cpp:199, quirk Q6)."""
    import subprocess
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    exe = tmp_path / "warp_partition_model"
    r = subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", "-Werror", "-o", str(exe), str(root / "tests" / "cpp" / "warp_partition_model.cpp")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe), "20000"], capture_output=True, text=True)
    assert r.returncode == 0 and "20000 ranges identical" in r.stdout, (r.stdout, r.stderr)


def test_reduction_order_of_the_absent_libraries_does_not_move_the_result(orc):
    """Eigen unrolls small fixed-size reductions as a balanced tree (a0 + (a1 + a2)) or runs them left to right depending on the
    expression and the build; the library is not in the image, so which one the reference's binary uses cannot be observed.  Measured
    instead: over easy and hard pairs the two orders give the same verdict, the same inlier lists and poses that agree to 1e-6, far inside
    the north star's 1e-5.  (The NUMBER of loop iterations may differ — an error that moves by one ulp flips a `refinedError <= rmse` and
    with it a skip-ahead — which is why the iteration trace is a parity item against the oracle only, not a portable property.)"""
    import synth
    worst = 0.0
    try:
        for seed, outl in ((42, 0.3), (7, 0.6), (9, 1.0), (21, 0.6), (22, 0.75), (24, 0.3), (1234, 0.05), (5, 0.45)):
            src, dst, m, _, _ = synth.rigid_pairs(seed=seed, outlier_frac=outl)
            orc.set_sum_order(False)
            a = orc.ransac_iterate(src, dst, m, seed=seed, depth_cov=1.6e-3)
            orc.set_sum_order(True)
            b = orc.ransac_iterate(src, dst, m, seed=seed, depth_cov=1.6e-3)
            assert a["ok"] == b["ok"]
            assert a["inliers"].tobytes() == b["inliers"].tobytes()
            worst = max(worst, float(np.abs(a["T12"] - b["T12"]).max()))
    finally:
        orc.set_sum_order(False)
    assert worst <= 1e-6
