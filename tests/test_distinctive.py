"""Landmark::ComputeDistinctiveDescriptors (Core/landmark.cpp:219-273): per landmark the observed
descriptor with the least median Hamming distance to the others.  Oracle vs a direct numpy statement (CPU), CUDA vs oracle."""
import numpy as np
import pytest

import synth


def make_landmarks(seed, n_landmarks, max_obs):
    rng = np.random.default_rng(seed)
    descs, offsets = [], [0]
    for l in range(n_landmarks):
        n = int(rng.integers(0 if l % 17 == 5 else 1, max_obs + 1))
        base = rng.integers(0, 256, 32, dtype=np.uint8)
        for _ in range(n):                                            # noisy observations of one descriptor, a few outliers
            bits = np.unpackbits(base)
            flip = rng.random(256) < (0.35 if rng.random() < 0.15 else 0.06)
            descs.append(np.packbits(bits ^ flip))
        offsets.append(offsets[-1] + n)
    d = np.stack(descs) if descs else np.zeros((0, 32), np.uint8)
    return d, np.array(offsets, np.int32)


def numpy_reference(desc, offsets):
    best = []
    for l in range(len(offsets) - 1):
        d = desc[offsets[l]:offsets[l + 1]]
        n = len(d)
        if n == 0:
            best.append(-1); continue
        D = np.unpackbits(d[:, None, :] ^ d[None, :, :], axis=-1).sum(-1).astype(np.float64)
        med = np.sort(D, axis=1)[:, int(0.5 * (n - 1))]
        best.append(int(np.argmin(med)))                              # first strictly smallest
    return np.array(best, np.int32)


def test_oracle_distinctive_descriptors(orc):
    desc, off = make_landmarks(1, 200, 24)
    best, med = orc.distinctive_descriptors(desc, off)
    assert np.array_equal(best, numpy_reference(desc, off))
    # a landmark whose observations are identical: median 0, first row wins
    same = np.repeat(desc[:1], 5, axis=0)
    b, m = orc.distinctive_descriptors(same, np.array([0, 5], np.int32))
    assert b[0] == 0 and m[0] == 0


@pytest.mark.gpu
def test_cuda_distinctive_descriptors(ob, orc):
    ctx = ob.Context(max_frames=1)
    try:
        for seed, n, mx in ((1, 200, 24), (2, 3000, 12), (3, 40, 100)):
            desc, off = make_landmarks(seed, n, mx)
            best, med = ctx.distinctive_descriptors(desc, off)
            bo, mo = orc.distinctive_descriptors(desc, off)
            assert np.array_equal(best, bo) and np.array_equal(med, mo), (seed, n, mx)
        b, m = ctx.distinctive_descriptors(np.zeros((0, 32), np.uint8), np.array([0, 0, 0], np.int32))
        assert list(b) == [-1, -1]
    finally:
        ctx.close()
