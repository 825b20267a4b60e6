"""Matcher::Fuse (search part, reference Features/matcher.cpp:212-296) and Matcher::BoWMatch (:145-209), SURVEY.md §8f rank 1:
oracle restatements pinned against independent numpy / cv2 replays, CUDA (orbf_fuse_search / orbf_bow_match) against the
oracle.  Bit-exact (indices, integer distances)."""
import numpy as np
import pytest

CAM = np.array([517.3, 516.5, 318.6, 255.3, 40.0, 0.0, 640.0, 0.0, 480.0], np.float32)      # FR1 intrinsics (common.h:35-38), mbf, image bounds


def _fuse_scene(seed, n_feat=1000, n_lm=800, radius=3.0, mono_frac=0.3):
    rng = np.random.default_rng(seed)
    ang = rng.normal(0, 0.05, 3)
    cx, sx, cy, sy, cz, sz = np.cos(ang[0]), np.sin(ang[0]), np.cos(ang[1]), np.sin(ang[1]), np.cos(ang[2]), np.sin(ang[2])
    R = (np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]]) @ np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]]) @ np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])).astype(np.float32)
    t = rng.normal(0, 0.05, 3).astype(np.float32)
    kp_x = rng.uniform(20, 620, n_feat).astype(np.float32); kp_y = rng.uniform(20, 460, n_feat).astype(np.float32)
    z = rng.uniform(0.8, 4.0, n_feat).astype(np.float32)
    u_right = (kp_x - CAM[4] / z).astype(np.float32)
    u_right[rng.random(n_feat) < mono_frac] = -1.0
    desc = rng.integers(0, 256, (n_feat, 32), dtype=np.uint8)
    src = rng.integers(0, n_feat, n_lm)
    # camera-frame point that projects near feature src, then moved to the world frame with the inverse pose
    jit = rng.normal(0, radius / 2.5, (n_lm, 2))
    pc = np.stack([((kp_x[src] + jit[:, 0]) - CAM[2]) * z[src] / CAM[0], ((kp_y[src] + jit[:, 1]) - CAM[3]) * z[src] / CAM[1], z[src]], 1).astype(np.float64)
    behind = rng.random(n_lm) < 0.05
    pc[behind, 2] *= -1
    pw = ((pc - t.astype(np.float64)) @ R.astype(np.float64)).astype(np.float32)           # R^T (pc - t)
    lm_desc = desc[src].copy()
    nb = rng.integers(0, 70, n_lm)
    for i in range(n_lm):
        bits = rng.choice(256, nb[i], replace=False)
        np.bitwise_xor.at(lm_desc[i], bits // 8, (1 << (bits % 8)).astype(np.uint8))
    valid = (rng.random(n_lm) < 0.9).astype(np.uint8)
    return R, t, kp_x, kp_y, u_right, desc, pw, lm_desc, valid


def _fuse_numpy(R, t, kp_x, kp_y, u_right, desc, pw, lm_desc, valid, radius, th_low):
    import cv2
    f32 = np.float32
    best = np.full(len(valid), -1, np.int32); dist = np.full(len(valid), -1, np.int32)
    fx, fy, cx, cy, mbf, x0, x1, y0, y1 = CAM
    for i in range(len(valid)):
        if not valid[i]:
            continue
        pc = cv2.gemm(R, pw[i].reshape(3, 1), 1.0, t.reshape(3, 1), 1.0)[:, 0]            # Rcw * p3Dw + tcw as cv::Mat evaluates it
        if pc[2] < 0:
            continue
        invz = f32(1) / pc[2]
        u = f32(fx * f32(pc[0] * invz)) + cx; v = f32(fy * f32(pc[1] * invz)) + cy
        if not (u >= x0 and u < x1 and v >= y0 and v < y1):
            continue
        ur = u - f32(mbf * invz)
        bd, bi = float("inf"), -1
        for j in np.nonzero((np.abs(kp_x - u) < f32(radius)) & (np.abs(kp_y - v) < f32(radius)))[0]:
            ex, ey = u - kp_x[j], v - kp_y[j]
            if u_right[j] >= 0:
                er = ur - u_right[j]
                if f32(f32(f32(ex * ex) + f32(ey * ey)) + f32(er * er)) > f32(7.8):
                    continue
            elif f32(f32(ex * ex) + f32(ey * ey)) > f32(5.99):
                continue
            d = float(np.unpackbits(lm_desc[i] ^ desc[j]).sum())
            if d < bd:
                bd, bi = d, j
        if bd <= th_low:
            best[i], dist[i] = bi, int(bd)
    return best, dist


def _bow_scene(seed, n1=900, n2=950, n_words=300, shared=0.7):
    rng = np.random.default_rng(seed)
    desc2 = rng.integers(0, 256, (n2, 32), dtype=np.uint8)
    src = rng.integers(0, n2, n1)
    desc1 = desc2[src].copy()
    nb = rng.integers(0, 60, n1)
    for i in range(n1):
        bits = rng.choice(256, nb[i], replace=False)
        np.bitwise_xor.at(desc1[i], bits // 8, (1 << (bits % 8)).astype(np.uint8))
    word2 = rng.integers(0, n_words, n2) * 3
    word1 = np.where(rng.random(n1) < shared, word2[src], rng.integers(0, n_words, n1) * 3 + 1)

    def csr(words, n):
        order = rng.permutation(n)                                   # bucket order is the insertion order, not sorted
        ids = np.unique(words)
        off = [0]; idx = []
        for w in ids:
            members = [int(k) for k in order if words[k] == w]
            idx += members; off.append(len(idx))
        return ids.astype(np.int32), np.array(off, np.int32), np.array(idx, np.int32)

    return (*csr(word1, n1), desc1, *csr(word2, n2), desc2)


def _bow_numpy(w1, o1, i1, d1, w2, o2, i2, d2, ratio, th_low):
    out = []; used = set()
    pos2 = {int(w): b for b, w in enumerate(w2)}
    for a, w in enumerate(w1):
        b = pos2.get(int(w))
        if b is None:
            continue
        for q in i1[o1[a]:o1[a + 1]]:
            b1 = b2 = float("inf"); bt = -1
            for t in i2[o2[b]:o2[b + 1]]:
                d = float(np.unpackbits(d1[q] ^ d2[t]).sum())
                if d < b1:
                    b2, b1, bt = b1, d, int(t)
                elif d < b2:
                    b2 = d
            if b1 <= th_low and np.float32(b1) < np.float32(ratio) * np.float32(b2) and bt not in used:
                out.append((int(q), bt, b1)); used.add(bt)
    return out


def test_fuse_oracle_matches_numpy_cv2(orc):
    sc = _fuse_scene(1, n_feat=300, n_lm=260)
    best, dist = orc.fuse_search(sc[0], sc[1], CAM, *sc[2:6], *sc[6:9], radius=3.0, th_low=50.0)
    rb, rd = _fuse_numpy(*sc, 3.0, 50.0)
    assert np.array_equal(best, rb) and np.array_equal(dist, rd)
    assert (best >= 0).sum() > 40


def test_bow_oracle_matches_numpy(orc):
    sc = _bow_scene(2, n1=250, n2=260, n_words=60)
    m = orc.bow_match(*sc, nn_ratio=0.6, th_low=50.0)
    ref = _bow_numpy(*sc, 0.6, 50.0)
    assert [(int(a), int(b), float(c)) for a, b, c in zip(m["queryIdx"], m["trainIdx"], m["distance"])] == ref
    assert (m["imgIdx"] == -1).all() and len(ref) > 30
    assert len(set(m["trainIdx"].tolist())) == len(m)                 # the std::set rule


@pytest.mark.gpu
@pytest.mark.parametrize("seed,n_feat,n_lm,radius", [(11, 1000, 800, 3.0), (12, 1000, 1500, 8.0), (13, 40, 7, 3.0), (14, 2000, 2000, 5.0)])
def test_fuse_cuda_matches_oracle(ob, orc, seed, n_feat, n_lm, radius):
    sc = _fuse_scene(seed, n_feat, n_lm, radius)
    rb, rd = orc.fuse_search(sc[0], sc[1], CAM, *sc[2:6], *sc[6:9], radius=radius, th_low=50.0)
    ctx = ob.Context(max_frames=2)
    gb, gd = ctx.fuse_search(sc[0], sc[1], CAM, sc[6], sc[7], sc[8], kp_x=sc[2], kp_y=sc[3], u_right=sc[4], desc=sc[5], radius=radius, th_low=50)
    assert np.array_equal(gb, rb) and np.array_equal(gd, rd)


@pytest.mark.gpu
def test_fuse_cuda_edge_cases(ob):
    sc = _fuse_scene(21, 64, 12)
    ctx = ob.Context(max_frames=2)
    gb, gd = ctx.fuse_search(sc[0], sc[1], CAM, sc[6][:0], sc[7][:0], sc[8][:0], kp_x=sc[2], kp_y=sc[3], u_right=sc[4], desc=sc[5])
    assert len(gb) == 0
    gb, gd = ctx.fuse_search(sc[0], sc[1], CAM, sc[6], sc[7], sc[8], kp_x=sc[2][:0], kp_y=sc[3][:0], u_right=sc[4][:0], desc=sc[5][:0])
    assert (gb == -1).all() and (gd == -1).all()
    gb, gd = ctx.fuse_search(sc[0], sc[1], CAM, sc[6], sc[7], np.zeros(12, np.uint8), kp_x=sc[2], kp_y=sc[3], u_right=sc[4], desc=sc[5])
    assert (gb == -1).all()


@pytest.mark.gpu
@pytest.mark.parametrize("seed,n1,n2,n_words", [(31, 900, 950, 300), (32, 1000, 1000, 40), (33, 30, 25, 5), (34, 2000, 2000, 1000)])
def test_bow_cuda_matches_oracle(ob, orc, seed, n1, n2, n_words):
    sc = _bow_scene(seed, n1, n2, n_words)
    ref = orc.bow_match(*sc, nn_ratio=0.6, th_low=50.0)
    ctx = ob.Context(max_frames=2)
    got = ctx.bow_match(*sc, nn_ratio=0.6, th_low=50)
    assert len(got) == len(ref)
    for f in ("queryIdx", "trainIdx", "imgIdx", "distance"):
        assert np.array_equal(got[f], ref[f]), f


@pytest.mark.gpu
def test_bow_cuda_edge_cases(ob, orc):
    sc = _bow_scene(41, 50, 50, 10)
    ctx = ob.Context(max_frames=2)
    e = np.zeros(0, np.int32)
    assert len(ctx.bow_match(e, np.zeros(1, np.int32), e, sc[3], *sc[4:])) == 0          # empty feature vector
    w1 = sc[0] + 100000                                                                  # no common word
    assert len(ctx.bow_match(w1, *sc[1:])) == 0
    assert len(orc.bow_match(w1, *sc[1:])) == 0
