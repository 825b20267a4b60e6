"""FAST edge cases of the round-2 kernel (csrc/fast.cu) against the oracle, through the C ABI:
  * thresholds >= 128 never enter the byte-parallel pretest (the carry test is written for th < 128): the whole strip takes the dense path;
  * the mask-free carry test lets a byte with |r - v| >= 129 + th carry into its neighbour — such images (hard black / white edges next
    to pixels whose difference is exactly th) must give the same corners: a carry may only add a pretest false positive;
  * survivor-list overflow (checkerboards, noise at the minimum threshold) falls back to the dense path strip by strip;
  * odd geometries: thresholds 1 / 127, saturated planes, isolated single-pixel spikes on the cell borders.
Candidates (incl. order), keypoints and descriptors are compared byte for byte (reference: cv::FAST per cell, orbextractor.cpp:665-723)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _same(ob, orc, img, **cfg):
    okw = {k: v for k, v in (("ini_th", cfg.get("ini_th_fast")), ("min_th", cfg.get("min_th_fast"))) if v is not None}
    ctx = ob.Context(max_frames=1, **cfg)
    try:
        ctx.extract_batch(img[None])
        k, d, _ = ctx.download_frame(0)
        ko, do, dbg = orc.extract(img, debug=True, **okw)
        off = 0
        for l in range(ctx.L):
            n = int(dbg["n_cands"][l])
            assert np.array_equal(ctx.level_candidates(0, l), dbg["cands"][off:off + n]), f"level {l}: candidates"
            off += n
        assert k.tobytes() == ko.tobytes() and np.array_equal(d, do)
        return len(k)
    finally:
        ctx.close()


def _texture(seed):
    import synth
    return synth.make_frame(synth.make_texture(seed, 480, 640), 3, 640, 480, seed)


@pytest.mark.parametrize("ini,mn", [(128, 100), (200, 150), (254, 128), (127, 7), (1, 1), (20, 20)])
def test_threshold_extremes(ob, orc, ini, mn):
    rng = np.random.default_rng(ini)
    img = _texture(2)
    img[100:300, 100:400] = rng.integers(0, 2, (200, 300), dtype=np.uint8) * 255      # a region with corners strong enough for any threshold
    _same(ob, orc, img, ini_th_fast=ini, min_th_fast=mn)


def test_carry_between_bytes_only_adds_false_positives(ob, orc):
    """Columns alternate between saturated differences (|r - v| = 255 >= 129 + th) and differences of exactly th."""
    img = np.full((480, 640), 100, np.uint8)
    img[:, 1::4] = 120                                   # |diff| == 20 == th against the 100 background
    img[:, 2::4] = 255
    img[::7, 3::4] = 0
    rng = np.random.default_rng(1)
    ys, xs = rng.integers(20, 460, 400), rng.integers(20, 620, 400)
    img[ys, xs] = rng.integers(0, 256, 400).astype(np.uint8)
    _same(ob, orc, img)


@pytest.mark.parametrize("period", [1, 2, 3, 5])
def test_checkerboards_overflow_the_survivor_list(ob, orc, period):
    yy, xx = np.mgrid[0:480, 0:640]
    img = ((((yy // period) + (xx // period)) & 1) * 255).astype(np.uint8)
    _same(ob, orc, img)


def test_noise_at_the_minimum_threshold(ob, orc):
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (480, 640), dtype=np.uint8)
    n = _same(ob, orc, img, ini_th_fast=7, min_th_fast=7)
    assert n >= 1000


def test_saturated_and_spiky_planes(ob, orc):
    for base in (0, 255):
        img = np.full((480, 640), base, np.uint8)
        assert _same(ob, orc, img) == 0
        img[16:464:35, 16:624:35] = 255 - base            # single-pixel spikes on the cell grid lines (NMS clipped to the cell, quirk Q1)
        img[19:464:35, 19:624:35] = 255 - base
        _same(ob, orc, img)


def test_low_contrast_everywhere_uses_the_fallback_threshold(ob, orc):
    img = (_texture(5).astype(np.float32) * 0.1 + 120).astype(np.uint8)       # nothing passes th = 20, cells are redone at th = 7
    _same(ob, orc, img)
