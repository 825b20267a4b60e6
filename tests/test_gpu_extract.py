"""GPU parity: every extraction stage of liborbfront_b200.so against the CPU oracle, through the C ABI.
Bars (BASELINE.json north_star): pyramid levels, FAST candidate lists (incl. order), quadtree keypoints
(incl. order), blurred levels, keypoint records, descriptors and 3D points are bit-exact; orientation is
compared bit-exact too (it is the same f32 polynomial), with the 1e-3 rad tolerance as the stated bar."""
import numpy as np
import pytest

import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx(ob):
    c = ob.Context(max_frames=4)
    yield c
    c.close()


def _check_frame(ctx, orc, img, depth=None, slot=0):
    kps_o, desc_o, dbg = orc.extract(img, debug=True)
    levels_o = orc.split_levels(dbg["pyramid"], dbg["ws"], dbg["hs"])
    blur_o = orc.split_levels(dbg["blurred"], dbg["ws"], dbg["hs"])
    off = 0
    for l in range(ctx.L):
        got = ctx.pyramid_level(slot, l)
        assert np.array_equal(got, levels_o[l]), f"pyramid level {l} differs"
        cand = ctx.level_candidates(slot, l)
        n = int(dbg["n_cands"][l])
        ref = dbg["cands"][off:off + n]; off += n
        assert len(cand) == n, f"level {l}: {len(cand)} candidates vs {n}"
        assert np.array_equal(cand, ref), f"FAST candidates of level {l} differ"
        if dbg["n_kps"][l] > 0:
            assert np.array_equal(ctx.pyramid_level(slot, l, blurred=True), blur_o[l]), f"blurred level {l} differs"
    assert np.array_equal(ctx.level_keypoint_counts(slot), dbg["n_kps"])
    kps, desc, xyz = ctx.download_frame(slot)
    assert len(kps) == len(kps_o)
    for f in ("x", "y", "size", "response", "octave", "class_id"):
        assert np.array_equal(kps[f], kps_o[f]), f"keypoint field {f} differs"
    dang = np.abs(kps["angle"] - kps_o["angle"])
    dang = np.minimum(dang, 360.0 - dang)
    assert np.deg2rad(dang.max()) <= 1e-3                      # north_star tolerance
    assert np.array_equal(kps["angle"], kps_o["angle"])        # and in fact bit-exact
    assert np.array_equal(desc, desc_o), "descriptors differ"
    if depth is not None:
        xyz_o, _ = orc.unproject(kps_o, depth)
        assert np.array_equal(xyz, xyz_o), "mvKeys3Dc differs"
    return kps, desc


def test_single_frame_drop_in(ctx, orc, texture):
    """orbf_extract == ORBextractor::operator() on one frame (host in / host out)."""
    img = synth.make_frame(texture, 0)
    kps, desc = ctx.extract(img)
    kps_o, desc_o = orc.extract(img)
    assert len(kps) == len(kps_o) >= 1000
    assert kps.tobytes() == kps_o.tobytes()
    assert np.array_equal(desc, desc_o)
    _check_frame(ctx, orc, img)


def test_batch_stagewise_with_fallback_cells(ctx, orc, texture):
    """4 frames incl. the low-contrast band frame (th=7 fallback cells), with depth, every stage compared."""
    idx = [5, 6, 7, 8]
    frames = np.stack([synth.make_frame(texture, i) for i in idx])
    depths = np.stack([synth.make_depth(i) for i in idx])
    ctx.extract_batch(frames, depths)
    counts = ctx.frame_counts(4)
    for s in range(4):
        kps, _ = _check_frame(ctx, orc, frames[s], depths[s], slot=s)
        assert counts[s] == len(kps)


def test_empty_image_leaves_outputs_untouched(ctx):
    kps, desc = ctx.extract(np.zeros((0, 0), np.uint8))
    assert len(kps) == 0 and len(desc) == 0


def test_flat_image_yields_no_keypoints(ctx, orc):
    img = np.full((480, 640), 77, np.uint8)
    kps, desc = ctx.extract(img)
    kps_o, _ = orc.extract(img)
    assert len(kps) == len(kps_o) == 0


def test_sparse_image_fewer_candidates_than_features(ctx, orc):
    """A handful of isolated corners: every level returns all its candidates (quadtree never reaches N)."""
    rng = np.random.default_rng(5)
    img = np.full((480, 640), 100, np.uint8)
    for _ in range(60):
        x, y = int(rng.integers(40, 600)), int(rng.integers(40, 440))
        img[y:y + 9, x:x + 9] = 220
    ctx.extract_batch(img[None])
    _check_frame(ctx, orc, img)


def test_random_noise_image_max_density(ctx, orc):
    """Uniform noise: far more candidates than the synthetic texture (stress for cell slots + quadtree)."""
    rng = np.random.default_rng(11)
    img = rng.integers(0, 256, (480, 640), dtype=np.uint8)
    ctx.extract_batch(img[None])
    _check_frame(ctx, orc, img)


def test_other_geometry_1280x720_2000(ob, orc):
    """BASELINE config 4 geometry: 1280x720, 2000 features (two quadtree roots)."""
    tex = synth.make_texture(3, 720, 1280)
    img = synth.make_frame(tex, 2, 1280, 720, seed=3)
    c = ob.Context(width=1280, height=720, nfeatures=2000, max_frames=1)
    try:
        kps, desc = c.extract(img)
        kps_o, desc_o = orc.extract(img, nfeatures=2000)
        assert kps.tobytes() == kps_o.tobytes()
        assert np.array_equal(desc, desc_o)
    finally:
        c.close()


def test_small_geometry_4_levels(ob, orc):
    tex = synth.make_texture(4, 240, 320)
    img = synth.make_frame(tex, 1, 320, 240, seed=4)
    c = ob.Context(width=320, height=240, nfeatures=300, nlevels=4, max_frames=1)
    try:
        kps, desc = c.extract(img)
        kps_o, desc_o = orc.extract(img, nfeatures=300, nlevels=4)
        assert kps.tobytes() == kps_o.tobytes()
        assert np.array_equal(desc, desc_o)
    finally:
        c.close()


def test_tables_match_oracle(ctx, orc):
    t = ctx.tables(); o = orc.tables()
    for k in ("scale", "inv_scale", "sigma2", "inv_sigma2", "nfeat"):
        assert np.array_equal(t[k], o[k]), k
    ws, hs = orc.level_sizes(640, 480)
    assert np.array_equal(t["level_w"], ws) and np.array_equal(t["level_h"], hs)
