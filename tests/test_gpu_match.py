"""GPU parity: Hamming kNN-2 + ratio (+ cross-check) through the C ABI vs the oracle (bit-exact distances,
indices and survivor lists; tie order = trainIdx ascending)."""
import numpy as np
import pytest

import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx(ob):
    c = ob.Context(max_frames=4)
    yield c
    c.close()


@pytest.mark.parametrize("maker", [synth.descriptor_sets, synth.tie_heavy_sets])
def test_knn2_bit_exact(ctx, orc, maker):
    A, B = maker()
    got = ctx.knn2(A, B)
    ref = orc.knn2(A, B)
    for g, r, name in zip(got, ref, ("idx1", "d1", "idx2", "d2")):
        assert np.array_equal(g, r), name


@pytest.mark.parametrize("ratio", [0.6, 0.8, 0.9])
@pytest.mark.parametrize("cross", [False, True])
def test_knn_match_ratio_cross(ctx, orc, ratio, cross):
    for maker in (synth.descriptor_sets, synth.tie_heavy_sets):
        A, B = maker()
        got = ctx.knn_match(A, B, ratio, cross)
        ref = orc.knn_match(A, B, ratio, cross)
        assert got.tobytes() == ref.tobytes()


def test_ragged_and_tiny_sets(ctx, orc):
    rng = np.random.default_rng(9)
    for nq, nt in ((1, 2), (5, 3), (33, 129), (1000, 2), (257, 1031), (1, 1), (7, 0), (0, 9)):
        A = rng.integers(0, 256, (nq, 32), dtype=np.uint8); B = rng.integers(0, 256, (nt, 32), dtype=np.uint8)
        got = ctx.knn2(A, B); ref = orc.knn2(A, B)
        for g, r in zip(got, ref):
            assert np.array_equal(g, r), (nq, nt)
        for cross in (False, True):
            assert ctx.knn_match(A, B, 0.8, cross).tobytes() == orc.knn_match(A, B, 0.8, cross).tobytes(), (nq, nt, cross)


def test_identical_sets_distance_zero(ctx, orc):
    A, _ = synth.descriptor_sets(300)
    i1, d1, i2, d2 = ctx.knn2(A, A)
    assert np.array_equal(i1, np.arange(300)) and (d1 == 0).all()
    assert ctx.knn_match(A, A, 0.9, True).tobytes() == orc.knn_match(A, A, 0.9, True).tobytes()


def test_symmetry_property_full_size(ctx):
    """Size-independent property at the full 1000x1000 size: d(q_i, t_idx1) recomputed on the host equals d1,
    and matching B->A gives distances consistent with A->B (Hamming is symmetric)."""
    A, B = synth.descriptor_sets(1000, seed=21)
    i1, d1, i2, d2 = ctx.knn2(A, B)
    dd = np.unpackbits(A ^ B[i1], axis=1).sum(axis=1)
    assert np.array_equal(dd, d1)
    assert (d1 <= d2).all()
    j1, e1, _, _ = ctx.knn2(B, A)
    assert (e1[i1] <= d1).all()


def test_descriptor_distance_host_helper(ob):
    a = np.arange(32, dtype=np.uint8); b = np.zeros(32, np.uint8)
    assert ob.descriptor_distance(a, b) == int(np.unpackbits(a).sum())
