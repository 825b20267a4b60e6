"""Matcher::ProjectionMatch (reference Features/matcher.cpp:90-143, Frame::GetFeaturesInArea Core/frame.cpp:258-274; SURVEY.md §8f
rank 1): the CUDA path (orbf_projection_match) against the oracle restatement, and the oracle against an independent numpy
replay of the reference loop.  Bit-exact (indices)."""
import numpy as np
import pytest

import synth


def _scene(seed, n_feat=1000, n_lm=900, radius=8.0, noise_bits=20, crowded=False):
    rng = np.random.default_rng(seed)
    w, h = (160, 120) if crowded else (640, 480)
    kp_x = rng.uniform(16, w - 16, n_feat).astype(np.float32)
    kp_y = rng.uniform(16, h - 16, n_feat).astype(np.float32)
    kp_oct = rng.integers(0, 8 if not crowded else 2, n_feat).astype(np.int32)
    desc = rng.integers(0, 256, (n_feat, 32), dtype=np.uint8)
    src = rng.integers(0, n_feat, n_lm)
    lm_desc = desc[src].copy()
    flip = rng.integers(0, 256, (n_lm, noise_bits))
    for k in range(noise_bits):
        lm_desc[np.arange(n_lm), flip[:, k] // 8] ^= (1 << (flip[:, k] % 8)).astype(np.uint8)
    proj_x = (kp_x[src] + rng.normal(0, radius / 3, n_lm)).astype(np.float32)
    proj_y = (kp_y[src] + rng.normal(0, radius / 3, n_lm)).astype(np.float32)
    flags = (rng.random(n_lm) < 0.9).astype(np.uint8) | ((rng.random(n_lm) < 0.8).astype(np.uint8) << 1)
    taken = (rng.random(n_feat) < 0.1).astype(np.uint8)
    return kp_x, kp_y, kp_oct, desc, lm_desc, proj_x, proj_y, flags, taken


def _numpy_replay(kp_x, kp_y, kp_oct, desc, lm_desc, proj_x, proj_y, flags, taken, radius, ratio, th_high):
    taken = taken.copy().astype(bool)
    best = np.full(len(flags), -1, np.int32)
    for i in range(len(flags)):
        if not flags[i] & 1:
            continue
        idx = np.nonzero((np.abs(kp_x - proj_x[i]) < np.float32(radius)) & (np.abs(kp_y - proj_y[i]) < np.float32(radius)))[0]
        if len(idx) == 0:
            continue
        b1 = b2 = float("inf"); l1 = l2 = -1; bi = -1
        for j in idx:
            if taken[j]:
                continue
            d = float(np.unpackbits(lm_desc[i] ^ desc[j]).sum())
            if d < b1:
                b2, b1, l2, l1, bi = b1, d, l1, kp_oct[j], j
            elif d < b2:
                l2, b2 = kp_oct[j], d
        if b1 <= th_high:
            if l1 == l2 and b1 > float(np.float32(ratio)) * b2:
                continue
            best[i] = bi
            if flags[i] & 2:
                taken[bi] = True
    return best


@pytest.mark.parametrize("seed,crowded", [(1, False), (2, True)])
def test_oracle_matches_numpy_replay(orc, seed, crowded):
    sc = _scene(seed, n_feat=300, n_lm=260, crowded=crowded)
    best, nm = orc.projection_match(*sc[:8], feat_taken=sc[8], radius=8.0, nn_ratio=0.8, th_high=100.0)
    ref = _numpy_replay(*sc, 8.0, 0.8, 100.0)
    assert np.array_equal(best, ref)
    assert nm == int((ref >= 0).sum())
    if crowded:
        assert (np.bincount(ref[ref >= 0]) > 1).any() or (ref >= 0).sum() > 50      # the scene does exercise contention


def test_oracle_edge_cases(orc):
    sc = _scene(3, n_feat=50, n_lm=10)
    best, nm = orc.projection_match(*sc[:4], sc[4][:0], sc[5][:0], sc[6][:0], sc[7][:0])
    assert len(best) == 0 and nm == 0
    flags = np.zeros(10, np.uint8)                                    # nothing in view
    best, nm = orc.projection_match(*sc[:7], flags)
    assert (best == -1).all() and nm == 0
    best, nm = orc.projection_match(*sc[:8], feat_taken=np.ones(50, np.uint8))   # every feature already taken
    assert (best == -1).all() and nm == 0


@pytest.mark.gpu
@pytest.mark.parametrize("seed,n_feat,n_lm,radius,crowded", [(11, 1000, 900, 8.0, False), (12, 1000, 1500, 15.0, False), (13, 400, 380, 8.0, True),
                                                            (14, 37, 5, 8.0, False), (15, 2000, 2000, 30.0, True),
                                                            (16, 50000, 300, 8.0, False), (17, 50000, 600, 1.5, False)])       # taken flags beyond the 48 KB default of shared memory: crowded / sparse windows
def test_cuda_matches_oracle_host_arrays(ob, orc, seed, n_feat, n_lm, radius, crowded):
    ctx = ob.Context(max_frames=2)
    sc = _scene(seed, n_feat, n_lm, radius, crowded=crowded)
    ref, nref = orc.projection_match(*sc[:8], feat_taken=sc[8], radius=radius, nn_ratio=0.8, th_high=100.0)
    got, ngot = ctx.projection_match(sc[4], sc[5], sc[6], sc[7], kp_x=sc[0], kp_y=sc[1], kp_octave=sc[2], desc=sc[3], feat_taken=sc[8], radius=radius,
                                     nn_ratio=0.8, th_high=100)
    assert np.array_equal(got, ref)
    assert ngot == nref


@pytest.mark.gpu
def test_cuda_on_a_frame_slot(ob, orc, texture):
    """The device-resident route: landmarks built from frame 0's own features, projected with a small offset into frame 1."""
    frames = np.stack([synth.make_frame(texture, i) for i in range(2)])
    depths = np.stack([synth.make_depth(i) for i in range(2)])
    ctx = ob.Context(max_frames=2)
    ctx.extract_batch(frames, depths)
    (k0, d0, _), (k1, d1, _) = ctx.download_frame(0), ctx.download_frame(1)
    rng = np.random.default_rng(5)
    n0 = len(k0)
    lm_desc = d0
    px = (k0["x"] + rng.normal(0, 2, n0)).astype(np.float32); py = (k0["y"] + rng.normal(0, 2, n0)).astype(np.float32)
    flags = np.full(n0, 3, np.uint8)
    ref, nref = orc.projection_match(k1["x"], k1["y"], k1["octave"], d1, lm_desc, px, py, flags, radius=8.0)
    got, ngot = ctx.projection_match(lm_desc, px, py, flags, slot=1, radius=8.0)
    assert np.array_equal(got, ref) and ngot == nref
    assert nref > 0


@pytest.mark.gpu
def test_cuda_edge_cases(ob):
    ctx = ob.Context(max_frames=2)
    sc = _scene(21, n_feat=64, n_lm=12)
    got, n = ctx.projection_match(sc[4][:0], sc[5][:0], sc[6][:0], sc[7][:0], kp_x=sc[0], kp_y=sc[1], kp_octave=sc[2], desc=sc[3])
    assert len(got) == 0 and n == 0
    got, n = ctx.projection_match(sc[4], sc[5], sc[6], sc[7], kp_x=sc[0][:0], kp_y=sc[1][:0], kp_octave=sc[2][:0], desc=sc[3][:0])
    assert (got == -1).all() and n == 0
    got, n = ctx.projection_match(sc[4], sc[5], sc[6], sc[7], kp_x=sc[0], kp_y=sc[1], kp_octave=sc[2], desc=sc[3], feat_taken=np.ones(64, np.uint8))
    assert (got == -1).all() and n == 0
