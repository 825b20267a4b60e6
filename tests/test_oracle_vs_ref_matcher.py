"""The oracle's restatements of Features/matcher.cpp — both KnnMatch overloads (row a-16), ProjectionMatch, BoWMatch and the search of
Fuse — against THE REFERENCE'S OWN SOURCE: matcher.cpp compiled verbatim into oracle/_ref/liborb_ref.so over
stand-in Frame / KeyFrame / Landmark classes (cv::BFMatcher::knnMatch, cv::norm and the cv::Mat product are the cv2-pinned oracle
routines).  Indices, distances and orders must be identical.

CPU-only.  Skipped where neither the reference checkout nor a prebuilt oracle/_ref exists."""
import numpy as np
import pytest

import synth
from test_fuse_bow import CAM, _bow_scene, _fuse_scene
from test_projection_match import _scene as _proj_scene


@pytest.fixture(scope="module")
def ref():
    from oracle import ref as r
    if not r.available():
        pytest.skip("oracle/_ref not built and /root/reference absent")
    return r


@pytest.mark.parametrize("ratio", [0.8, 0.6, 0.95])
def test_knn_match_frames(ref, orc, ratio):
    for seed in (1, 2):
        q, t = synth.descriptor_sets(n=400, seed=seed)[:2]
        assert ref.knn_match_frames(q, t, ratio).tobytes() == orc.knn_match(q, t, ratio, False).tobytes()
    q, t = synth.tie_heavy_sets(n=300, seed=3)[:2]                     # equal distances: (distance, trainIdx) order decides
    assert ref.knn_match_frames(q, t, ratio).tobytes() == orc.knn_match(q, t, ratio, False).tobytes()
    assert len(ref.knn_match_frames(q[:0], t, ratio)) == 0


def test_knn_match_keyframe_landmark_filters(ref, orc):
    rng = np.random.default_rng(5)
    q, t = synth.descriptor_sets(n=500, seed=7)[:2]
    n_ids = 600
    kf_lm = rng.integers(1, n_ids, len(q)).astype(np.int32); kf_lm[rng.random(len(q)) < 0.15] = 0
    bad = (rng.random(n_ids) < 0.1).astype(np.uint8)
    f2_a = np.zeros(len(t), np.int32); f2_a[rng.random(len(t)) < 0.2] = 7
    f2_b = f2_a.astype(np.int64)
    got = ref.knn_match_keyframe(q, t, 0.8, kf_lm, bad, f2_a)
    exp = orc.knn_match_keyframe(q, t, 0.8, kf_lm, lambda p: bool(bad[p]), f2_b)
    assert got.tobytes() == exp.tobytes() and len(got) > 50
    assert np.array_equal(f2_a, f2_b.astype(np.int32))                 # the landmarks handed to the frame


@pytest.mark.parametrize("seed,crowded", [(1, False), (2, True), (3, True)])
def test_projection_match(ref, orc, seed, crowded):
    sc = _proj_scene(seed, n_feat=400, n_lm=380, crowded=crowded)
    for ratio, th_high in ((0.8, 100.0), (0.6, 60.0)):
        b_r, n_r = ref.projection_match(*sc[:8], feat_taken=sc[8], radius=8.0, nn_ratio=ratio, th_high=th_high)
        b_o, n_o = orc.projection_match(*sc[:8], feat_taken=sc[8], radius=8.0, nn_ratio=ratio, th_high=th_high)
        assert np.array_equal(b_r, b_o) and n_r == n_o and n_r > 20
    b_r, n_r = ref.projection_match(*sc[:8])                           # no slot taken at the start
    b_o, n_o = orc.projection_match(*sc[:8])
    assert np.array_equal(b_r, b_o) and n_r == n_o


@pytest.mark.parametrize("seed", [2, 9])
def test_bow_match(ref, orc, seed):
    sc = _bow_scene(seed, n1=300, n2=320, n_words=70)
    for ratio, th in ((0.6, 50.0), (0.9, 80.0)):
        m_r = ref.bow_match(*sc, nn_ratio=ratio, th_low=th); m_o = orc.bow_match(*sc, nn_ratio=ratio, th_low=th)
        assert m_r.tobytes() == m_o.tobytes() and len(m_r) > 20


@pytest.mark.parametrize("seed,radius", [(1, 3.0), (4, 8.0)])
def test_fuse_search(ref, orc, seed, radius):
    R, t, kp_x, kp_y, u_right, desc, pw, lm_desc, valid = _fuse_scene(seed, n_feat=400, n_lm=360, radius=radius)
    state = valid.copy()
    inval = np.nonzero(valid == 0)[0]
    state[inval] = np.array([0, 2, 3], np.uint8)[np.arange(len(inval)) % 3]   # null pointer / isBad() / IsInKeyFrame(pKF): all skipped
    b_r, nf = ref.fuse(R, t, CAM[5:9], kp_x, kp_y, u_right, desc, pw, lm_desc, state, radius=radius, th_low=50.0)
    b_o, _ = orc.fuse_search(R, t, CAM, kp_x, kp_y, u_right, desc, pw, lm_desc, valid, radius=radius, th_low=50.0)
    assert np.array_equal(b_r, b_o)
    assert nf == int((b_o >= 0).sum()) and nf > 30
