"""The oracle's restatement of rows a-11 ... a-15 (Ransac::Iterate, SampleMatches, GetTransformFromMatches, ComputeInliersAndError,
ErrorFunction2, Kabsch::Compute) against THE REFERENCE'S OWN SOURCE: oracle/_ref/libodometry_ref.so is
/root/reference/Odometry/ransac.cpp + kabsch.cpp compiled verbatim (oracle/Makefile, target _ref) against stand-ins for Eigen / PCL /
boost / OpenCV.  The three third-party numerical routines (PCL TransformationFromCorrespondences, Eigen 3x3 LLT solve, Jacobi SVD)
are one shared restatement (those libraries are not in the image), so every result must be bit-identical: the depth filter, the
std::sort order, the libc rand() sample loop, the refinement loop, the accept / skip-ahead / early-exit rule, the identity
fallback, the inlier rule with quirks Q7 / Q8, the clouds, the Kabsch flow — all run from the reference's source.

CPU-only.  Skipped where neither the reference checkout nor a prebuilt oracle/_ref exists."""
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

import synth

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def ref():
    from oracle import ref as r
    if not r.available():
        pytest.skip("oracle/_ref not built and /root/reference absent")
    return r


@pytest.fixture(scope="module")
def cov(ref):
    """Quirk Q7: the reference's depth covariance is a process-wide static fixed by the first depth it sees.  Latch it here with a
    chosen depth (2.0 m -> (0.01 * 4)^2) — or read what an earlier test of this process latched — and give the oracle the same value."""
    return ref.depth_covariance(2.0)


def _same(r, o, cov):
    assert r["ok"] == o["ok"]
    assert r["n_good"] == o["n_good"]
    assert r["inliers"].tobytes() == o["inliers"].tobytes(), "inlier lists differ"
    assert np.array_equal(r["T12"], o["T12"]), "T12 differs"
    assert r["rmse"] == o["rmse"]
    assert o["depth_cov"] == cov


@pytest.mark.parametrize("seed", [42, 7, 1234, 99])
@pytest.mark.parametrize("outliers", [0.05, 0.3, 0.6, 0.75])
def test_iterate_identical_to_reference_source(ref, orc, cov, seed, outliers):
    src, dst, m, _, _ = synth.rigid_pairs(seed=seed, outlier_frac=outliers)
    r = ref.ransac_iterate(src, dst, m, seed=seed)
    o = orc.ransac_iterate(src, dst, m, seed=seed, depth_cov=cov)
    _same(r, o, cov)
    if outliers <= 0.3:
        assert r["ok"] and len(r["inliers"]) >= 20


def test_member_form_is_the_same_loop(ref, orc, cov):
    """Ransac(KeyFrame*, KeyFrame*, matches) + Iterate() (ransac.cpp:26-36, 44-153) against the same oracle call."""
    for seed, outl in ((3, 0.3), (4, 0.7)):
        src, dst, m, _, _ = synth.rigid_pairs(seed=seed, outlier_frac=outl)
        r = ref.ransac_iterate(src, dst, m, seed=seed, member_form=True)
        _same(r, orc.ransac_iterate(src, dst, m, seed=seed, depth_cov=cov), cov)


def test_hard_cases_all_outliers_identity_and_too_few(ref, orc, cov):
    src, dst, m, _, _ = synth.rigid_pairs(seed=8, outlier_frac=1.0)                  # all 200 iterations, no model
    _same(ref.ransac_iterate(src, dst, m, seed=3), orc.ransac_iterate(src, dst, m, seed=3, depth_cov=cov), cov)
    src, _, m, _, _ = synth.rigid_pairs(seed=10, outlier_frac=0.0)                    # dst == src
    dst = np.zeros_like(src); dst[m["trainIdx"]] = src[m["queryIdx"]]
    _same(ref.ransac_iterate(src, dst, m, seed=2), orc.ransac_iterate(src, dst, m, seed=2, depth_cov=cov), cov)
    src, dst, m, _, _ = synth.rigid_pairs(seed=6)
    r = ref.ransac_iterate(src, dst, m[:10]); o = orc.ransac_iterate(src, dst, m[:10], depth_cov=cov)   # fewer than mMinInlierTh matches
    assert not r["ok"] and not o["ok"] and r["rmse"] == o["rmse"] == 1e6 and len(r["inliers"]) == 0
    assert np.array_equal(r["T12"], np.eye(4, dtype=np.float32)) and len(r["cloud_src"]) == 0


def test_identity_fallback_branch(ref, orc, cov):
    """validIters == 0 and the identity transform explains the matches (ransac.cpp:252-264): tiny motion, but every hypothesis
    is rejected because min_inlier_th exceeds what 4-point models... is forced by a sample size larger than the match list cannot be —
    so use sample_size > good matches: the loop never runs and only the fallback decides."""
    src, _, m, _, _ = synth.rigid_pairs(seed=11, outlier_frac=0.0)
    dst = np.zeros_like(src); dst[m["trainIdx"]] = src[m["queryIdx"]]
    m = m[:60]
    kw = dict(sample_size=100, min_inlier_th=20, seed=5)
    r = ref.ransac_iterate(src, dst, m, **kw); o = orc.ransac_iterate(src, dst, m, depth_cov=cov, **kw)
    assert o["used_identity"] and o["real_iters"] == 0
    _same(r, o, cov)
    assert r["ok"] and r["rmse"] == o["rmse"] and r["rmse"] >= 1e6        # rmse += inlierError on top of 1e6 (sic)


def test_other_parameters_and_no_depth_check(ref, orc, cov):
    src, dst, m, _, _ = synth.rigid_pairs(seed=13, outlier_frac=0.4)
    kw = dict(iterations=50, min_inlier_th=30, max_mahal=2.0, sample_size=3, seed=17)
    _same(ref.ransac_iterate(src, dst, m, **kw), orc.ransac_iterate(src, dst, m, depth_cov=cov, **kw), cov)
    # invalid depths (NaN, zero, negative) in the lists, with and without mCheckDepth
    src = src.copy(); dst = dst.copy()
    src[m["queryIdx"][::7], 2] = np.nan; dst[m["trainIdx"][3::11], 2] = 0.0; src[m["queryIdx"][5::13], 2] = -1.0
    dst[m["trainIdx"][2::17], 0] = 0.0                                     # quirk Q8: target.x == 0 is skipped by the inlier rule
    for chk in (True, False):
        r = ref.ransac_iterate(src, dst, m, seed=21, check_depth=chk)
        o = orc.ransac_iterate(src, dst, m, seed=21, check_depth=chk, depth_cov=cov)
        _same(r, o, cov)


def test_clouds_are_the_depth_filtered_matches_in_match_order(ref, orc, cov):
    src, dst, m, _, _ = synth.rigid_pairs(seed=15, outlier_frac=0.2)
    src = src.copy(); src[m["queryIdx"][::9], 2] = np.nan
    r = ref.ransac_iterate(src, dst, m, seed=1)
    cs, ct = orc.ransac_clouds(src, dst, m)
    assert np.array_equal(r["cloud_src"], cs[:, :3]) and np.array_equal(r["cloud_tgt"], ct[:, :3])


@pytest.mark.parametrize("M", [4, 5, 37, 400, 1000])
def test_sample_matches_tables(ref, orc, M):
    for seed in (1, 42, 2024):
        assert np.array_equal(ref.sample_table(seed, M), orc.sample_table_libc(seed, M))
    assert np.array_equal(ref.sample_table(9, M, 50, 3), orc.sample_table_libc(9, M, 50, 3)) or M < 3


def test_compute_inliers_and_error(ref, orc, cov):
    src, dst, m, R, t = synth.rigid_pairs(seed=17, outlier_frac=0.3)
    T = np.eye(4, dtype=np.float32); T[:3, :3] = R; T[:3, 3] = t
    for Tm in (T, np.eye(4, dtype=np.float32)):
        e, pos = ref.inliers_and_error(src, dst, m, Tm)
        d = np.array([orc.mahalanobis2(src[a["queryIdx"]], dst[a["trainIdx"]], Tm, cov) for a in m])
        skip = (src[m["queryIdx"], 2] == 0) | (dst[m["trainIdx"], 0] == 0)
        keep = ~skip & (d <= np.float32(3.0) * np.float32(3.0)) & (d >= 0)
        assert np.array_equal(pos, np.nonzero(keep)[0])
        if keep.sum() >= 3:
            acc = 0.0
            for v in d[keep]:
                acc += float(v)
            assert e == np.sqrt(acc / int(keep.sum()))
        else:
            assert e == 1e9


def test_kabsch_flow(ref, orc):
    src, dst, m, R, t = synth.rigid_pairs(seed=14, outlier_frac=0.0)
    A = src[m["queryIdx"]]; A = A[A[:, 2] > 0][:180]
    B = (A.astype(np.float64) @ R.T + t).astype(np.float32)
    assert np.array_equal(ref.kabsch(A, B), orc.kabsch(A, B))
    assert np.abs(ref.kabsch(A, B)[:3, :3] - R).max() < 1e-4
    Bm = A * np.array([1, 1, -1], np.float32)                             # reflection: det < 0 flips the last axis
    assert np.array_equal(ref.kabsch(A, Bm), orc.kabsch(A, Bm))
    assert np.array_equal(ref.kabsch(np.zeros((0, 3)), np.zeros((0, 3))), np.eye(4, dtype=np.float32))
    rng = np.random.default_rng(5)
    for n in (3, 4, 17):
        P = rng.normal(size=(n, 3)).astype(np.float32); Q = rng.normal(size=(n, 3)).astype(np.float32)
        assert np.array_equal(ref.kabsch(P, Q), orc.kabsch(P, Q))


_LATCH_SNIPPET = r"""
import sys, json
sys.path.insert(0, {root!r}); sys.path.insert(0, {root!r} + "/tests")
import numpy as np, synth
from oracle import ref, oracle as orc
src, dst, m, _, _ = synth.rigid_pairs(seed=31, outlier_frac=0.3)
r = ref.ransac_iterate(src, dst, m, seed=31)              # the first ErrorFunction2 call of this process latches the covariance
latched = ref.depth_covariance(123.0)                      # any later call returns the latched value
o = orc.ransac_iterate(src, dst, m, seed=31, depth_cov=-1.0)   # oracle rule: latch from the first pair that reaches scoring
print(json.dumps(dict(latched=latched, oracle=o["depth_cov"], same_inliers=r["inliers"].tobytes() == o["inliers"].tobytes(),
                      same_T=bool(np.array_equal(r["T12"], o["T12"])))))
"""


def test_depth_covariance_latch_rule_in_a_fresh_process(ref):
    """Quirk Q7 end to end: in a new process the reference latches the covariance of the first point it scores; the oracle's
    explicit rule (depth_cov < 0: latch from the first scored match) must produce the same number and the same result."""
    import json
    out = subprocess.run([sys.executable, "-c", _LATCH_SNIPPET.format(root=str(ROOT))], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads(out.stdout.strip().splitlines()[-1])
    assert d["latched"] == d["oracle"] and d["latched"] > 0
    assert d["same_inliers"] and d["same_T"]
