"""Golden-vector tests (tests/golden/, minted by tools/make_goldens.py from cv2 4.13.0 entry points + the oracle).

  not gpu : the CPU oracle reproduces every golden (pins the checker itself);
  gpu     : the CUDA path, called through the C ABI, reproduces every golden.
Inputs are regenerated from tests/synth.py seeds and checked against the sha256 stored with the golden, so a drift of
the generator cannot silently re-baseline anything.  Bar: bit-exact for pixels, candidates, keypoints, descriptors,
distances, indices and inlier sets; pose within 1e-5 (BASELINE.json north_star), written below."""
import hashlib
from pathlib import Path

import numpy as np
import pytest

import synth

GOLD = Path(__file__).resolve().parent / "golden"
POSE_TOL = 1e-5


def sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), np.uint8)


def split(buf, wh):
    out, o = [], 0
    for w, h in wh:
        out.append(buf[o:o + w * h].reshape(h, w)); o += w * h
    return out


@pytest.fixture(scope="module")
def g_small():
    return np.load(GOLD / "extract_320x240.npz")


@pytest.fixture(scope="module")
def g_full():
    return np.load(GOLD / "extract_640x480.npz")


@pytest.fixture(scope="module")
def g_match():
    return np.load(GOLD / "match_knn2.npz")


@pytest.fixture(scope="module")
def g_ransac():
    return np.load(GOLD / "ransac.npz")


def small_inputs(g):
    w, h, nfeat, nlev, seed = (int(v) for v in g["params"])
    tex = synth.make_texture(seed, h, w)
    out = []
    for i in g["frame_ids"]:
        img = synth.make_frame(tex, int(i), w, h, seed); depth = synth.make_depth(int(i), w, h, seed)
        assert np.array_equal(sha(img), g[f"f{i}_input_sha"]) and np.array_equal(sha(depth), g[f"f{i}_depth_sha"]), "synthetic input drifted"
        out.append((int(i), img, depth))
    return (w, h, nfeat, nlev), out


MATCH_SETS = {"rand": lambda: synth.descriptor_sets(1000, seed=1), "ties": lambda: synth.tie_heavy_sets(1000, seed=3),
              "ragged": lambda: tuple(x[:n] for x, n in zip(synth.descriptor_sets(700, seed=9), (613, 257))),
              "tiny": lambda: tuple(x[:n] for x, n in zip(synth.descriptor_sets(64, seed=11), (5, 2)))}
RANSAC_SETS = {"a": (4, 650), "b": (5, 300), "few": (6, 24)}


# ---------------------------------------------------------------------------------------------------------------------
# CPU: the oracle against the goldens
# ---------------------------------------------------------------------------------------------------------------------
def test_oracle_extraction_stages_match_goldens(orc, g_small):
    (w, h, nfeat, nlev), frames = small_inputs(g_small)
    for i, img, depth in frames:
        p = f"f{i}_"
        k, d, dbg = orc.extract(img, nfeatures=nfeat, nlevels=nlev, debug=True)
        assert np.array_equal(dbg["pyramid"], g_small[p + "pyramid"])                        # cv2.resize chain
        assert np.array_equal(dbg["n_cands"], g_small[p + "cand_counts"])                     # cv2.FastFeatureDetector per cell
        c = dbg["cands"]
        assert np.array_equal(np.stack([c["x"], c["y"], c["score"]], 1), g_small[p + "cands"])
        assert np.array_equal(dbg["n_kps"], g_small[p + "kp_counts"])
        wh = g_small[p + "level_wh"]
        for l, (a, b) in enumerate(zip(split(dbg["blurred"], wh), split(g_small[p + "blurred"], wh))):
            if dbg["n_kps"][l]:
                assert np.array_equal(a, b), f"blurred level {l}"                            # cv2.GaussianBlur
        assert k.tobytes() == g_small[p + "keypoints"].tobytes() and np.array_equal(d, g_small[p + "descriptors"])
        xyz, ur = orc.unproject(k, depth)
        assert np.array_equal(xyz, g_small[p + "xyz"]) and np.array_equal(ur, g_small[p + "uright"])


def test_oracle_full_size_matches_goldens(orc, g_full, texture):
    for i in g_full["frame_ids"]:
        img = synth.make_frame(texture, int(i))
        assert np.array_equal(sha(img), g_full[f"f{i}_input_sha"])
        k, d, dbg = orc.extract(img, debug=True)
        wh = list(zip(dbg["ws"], dbg["hs"]))
        assert all(np.array_equal(sha(a), s) for a, s in zip(split(dbg["pyramid"], wh), g_full[f"f{i}_pyramid_sha"]))
        assert all(np.array_equal(sha(a), s) for a, s in zip(split(dbg["blurred"], wh), g_full[f"f{i}_blurred_sha"]))
        assert np.array_equal(dbg["n_cands"], g_full[f"f{i}_cand_counts"]) and np.array_equal(dbg["n_kps"], g_full[f"f{i}_kp_counts"])
        assert k.tobytes() == g_full[f"f{i}_keypoints"].tobytes() and np.array_equal(d, g_full[f"f{i}_descriptors"])


@pytest.mark.parametrize("tag", list(MATCH_SETS))
def test_oracle_matching_matches_goldens(orc, g_match, tag):
    A, B = MATCH_SETS[tag]()
    assert np.array_equal(sha(A), g_match[tag + "_A_sha"]) and np.array_equal(sha(B), g_match[tag + "_B_sha"])
    assert np.array_equal(np.stack(orc.knn2(A, B), 1), g_match[tag + "_knn"])
    for r in (6, 8, 9):
        assert np.array_equal(orc.knn_match(A, B, r / 10)["queryIdx"], g_match[f"{tag}_ratio{r}"])


@pytest.mark.parametrize("tag", list(RANSAC_SETS))
def test_oracle_ransac_matches_goldens(orc, g_ransac, tag):
    seed, m = RANSAC_SETS[tag]
    src, dst, matches, _, _ = synth.rigid_pairs(m=m, seed=seed)
    r = orc.ransac_iterate(src, dst, matches, seed=42)
    assert r["good_sorted"].tobytes() == g_ransac[tag + "_good_sorted"].tobytes()
    assert np.array_equal(r["sample_table"], g_ransac[tag + "_sample_table"])
    assert r["inliers"].tobytes() == g_ransac[tag + "_inliers"].tobytes()
    assert np.abs(r["T12"] - g_ransac[tag + "_T12"]).max() <= POSE_TOL
    assert [int(r["ok"]), r["n_good"], r["real_iters"], r["valid_iters"], int(r["used_identity"])] == list(g_ransac[tag + "_scalars"])


def test_ransac_golden_recovers_the_seeded_motion(g_ransac):
    """Domain sanity of the golden itself: the accepted pose is the rigid motion the data was generated with."""
    for tag in ("a", "b"):
        T = g_ransac[tag + "_T12"]; truth = g_ransac[tag + "_truth_Rt"]
        assert np.abs(T[:3, :3] - truth[:9].reshape(3, 3)).max() < 5e-3 and np.abs(T[:3, 3] - truth[9:]).max() < 5e-3


# ---------------------------------------------------------------------------------------------------------------------
# GPU: the CUDA path (C ABI) against the goldens
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_cuda_extraction_stages_match_goldens(ob, g_small):
    (w, h, nfeat, nlev), frames = small_inputs(g_small)
    ctx = ob.Context(width=w, height=h, nfeatures=nfeat, nlevels=nlev, max_frames=len(frames))
    try:
        ctx.extract_batch(np.stack([f[1] for f in frames]), np.stack([f[2] for f in frames]))
        for s, (i, img, depth) in enumerate(frames):
            p = f"f{i}_"
            wh = g_small[p + "level_wh"]
            counts = ctx.level_keypoint_counts(s)
            assert np.array_equal(counts, g_small[p + "kp_counts"])
            gp, gb = split(g_small[p + "pyramid"], wh), split(g_small[p + "blurred"], wh)
            off = 0
            for l in range(nlev):
                assert np.array_equal(ctx.pyramid_level(s, l), gp[l]), f"pyramid level {l}"
                assert np.array_equal(ctx.pyramid_level(s, l, blurred=True), gb[l]), f"blurred level {l}"
                c = ctx.level_candidates(s, l)
                n = int(g_small[p + "cand_counts"][l])
                assert np.array_equal(np.stack([c["x"], c["y"], c["score"]], 1), g_small[p + "cands"][off:off + n]), f"candidates level {l}"
                off += n
            k, d, xyz = ctx.download_frame(s)
            assert k.tobytes() == g_small[p + "keypoints"].tobytes() and np.array_equal(d, g_small[p + "descriptors"])
            assert np.array_equal(xyz, g_small[p + "xyz"])
    finally:
        ctx.close()


@pytest.mark.gpu
def test_cuda_full_size_matches_goldens(ob, g_full, texture):
    ids = [int(i) for i in g_full["frame_ids"]]
    frames = np.stack([synth.make_frame(texture, i) for i in ids])
    ctx = ob.Context(max_frames=len(ids))
    try:
        ctx.extract_batch(frames)
        for s, i in enumerate(ids):
            for l in range(8):
                assert np.array_equal(sha(ctx.pyramid_level(s, l)), g_full[f"f{i}_pyramid_sha"][l])
                assert np.array_equal(sha(ctx.pyramid_level(s, l, blurred=True)), g_full[f"f{i}_blurred_sha"][l])
                assert len(ctx.level_candidates(s, l)) == g_full[f"f{i}_cand_counts"][l]
            k, d, _ = ctx.download_frame(s)
            assert k.tobytes() == g_full[f"f{i}_keypoints"].tobytes() and np.array_equal(d, g_full[f"f{i}_descriptors"])
            k1, d1 = ctx.extract(frames[s])                                     # ORBextractor::operator() single-frame entry
            assert k1.tobytes() == k.tobytes() and np.array_equal(d1, d)
    finally:
        ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("tag", list(MATCH_SETS))
def test_cuda_matching_matches_goldens(ob, g_match, tag):
    A, B = MATCH_SETS[tag]()
    ctx = ob.Context(max_frames=1)
    try:
        assert np.array_equal(np.stack(ctx.knn2(A, B), 1), g_match[tag + "_knn"])
        for r in (6, 8, 9):
            assert np.array_equal(ctx.knn_match(A, B, r / 10)["queryIdx"], g_match[f"{tag}_ratio{r}"])
        assert np.array_equal(ctx.knn_match(A, B, 0.8, cross_check=True)["queryIdx"], g_match[tag + "_cross8"])
    finally:
        ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("tag", list(RANSAC_SETS))
def test_cuda_ransac_matches_goldens(ob, g_ransac, tag):
    seed, m = RANSAC_SETS[tag]
    src, dst, matches, _, _ = synth.rigid_pairs(m=m, seed=seed)
    ctx = ob.Context(max_frames=1)
    try:
        r = ctx.ransac_iterate(src, dst, matches, seed=42)
        if r["n_good"] >= 20:    # below minInlierTh Ransac::Iterate returns before it sorts or samples (ransac.cpp:191-192)
            assert r["good_sorted"].tobytes() == g_ransac[tag + "_good_sorted"].tobytes()
            assert np.array_equal(r["sample_table"], g_ransac[tag + "_sample_table"])
        assert r["inliers"].tobytes() == g_ransac[tag + "_inliers"].tobytes()
        assert np.abs(r["T12"] - g_ransac[tag + "_T12"]).max() <= POSE_TOL
        assert [int(r["ok"]), r["n_good"], r["real_iters"], r["valid_iters"], int(r["used_identity"])] == list(g_ransac[tag + "_scalars"])
        n = r["real_iters"]      # the GPU scores whole waves of hypotheses; only the rows the sequential loop consumed are defined
        assert np.array_equal(r["hyp"]["n_refined"][:n], g_ransac[tag + "_hyp_n"][:n])
    finally:
        ctx.close()


@pytest.mark.gpu
def test_cuda_kabsch_matches_goldens(ob, g_ransac):
    rng = np.random.default_rng(8)
    A = rng.normal(size=(40, 3)).astype(np.float32)
    ang = 0.3
    Rz = np.array([[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]], np.float32)
    B = (A @ Rz.T + np.array([0.1, -0.2, 0.3], np.float32)).astype(np.float32)
    ctx = ob.Context(max_frames=1)
    try:
        assert np.abs(ctx.kabsch(A, B) - g_ransac["kabsch_T"]).max() <= POSE_TOL
        assert np.abs(ctx.kabsch(A, (A * np.array([1, 1, -1], np.float32)).astype(np.float32)) - g_ransac["kabsch_reflect_T"]).max() <= POSE_TOL
        assert np.array_equal(ctx.kabsch(np.zeros((0, 3), np.float32), np.zeros((0, 3), np.float32)), g_ransac["kabsch_empty_T"])
    finally:
        ctx.close()


def test_glibc_rand_replay_matches_golden(ob, g_ransac):
    """csrc/replay.h's glibc rand() restatement (host build of the same code the kernel runs) vs the real libc sequence."""
    assert np.array_equal(ob.selftest_glibc_rand(42, 64), g_ransac["libc_rand_seed42"])
