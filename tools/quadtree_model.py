"""Python model of csrc/quadtree.cu's round-based formulation; checked against the list-based oracle.
Run: python tools/quadtree_model.py   (authoring aid; tests/test_quadtree_model.py runs a short version)"""
import math
import sys
from pathlib import Path

import numpy as np

sys.path[:0] = [str(Path(__file__).resolve().parent.parent), str(Path(__file__).resolve().parent.parent / "tests")]


def quadrant(x, y, r):
    mx = r[0] + ((r[2] - r[0] + 1) >> 1); my = r[1] + ((r[3] - r[1] + 1) >> 1)
    return (0 if y < my else 2) if x < mx else (1 if y < my else 3)


def distribute_rounds(cands, W, H, N):
    """cands: list of (x, y, score) relative to minBorder. Returns candidate indices in list order."""
    n = len(cands)
    if n == 0:
        return []
    f32 = np.float32
    nIni = int(math.floor(float(f32(W) / f32(H)) + 0.5))
    hX = f32(W) / f32(nIni)
    rootcnt = [0] * nIni
    node_of = []
    for (x, y, s) in cands:
        r = min(int(f32(x) / hX), nIni - 1)
        node_of.append(r); rootcnt[r] += 1
    rect, cnt, pos_of_root = [], [], {}
    for r in range(nIni):
        if rootcnt[r] > 0:
            pos_of_root[r] = len(rect)
            rect.append((int(hX * f32(r)), 0, int(hX * f32(r + 1)), H)); cnt.append(rootcnt[r])
    node_of = [pos_of_root[r] for r in node_of]
    m, Efresh, phase2 = len(rect), 0, False
    while True:
        cc = [[0, 0, 0, 0] for _ in range(m)]
        def is_cand(i):
            return cnt[i] > 1 and (not phase2 or i < Efresh)
        for k, (x, y, s) in enumerate(cands):
            nd = node_of[k]
            if is_cand(nd):
                cc[nd][quadrant(x, y, rect[nd])] += 1
        cand_nodes = [i for i in range(m) if is_cand(i)]
        if not phase2:
            order = cand_nodes
        else:
            order = sorted(cand_nodes, key=lambda i: (-cnt[i], i))
            size = m; cut = len(order) - 1
            for r, i in enumerate(order):
                size += sum(1 for c in cc[i] if c > 0) - 1
                if size >= N:
                    cut = r; break
            order = order[:cut + 1]
        expd = set(order)
        created = []           # (parent, q) in creation order
        for i in order:
            for q in range(4):
                if cc[i][q] > 0:
                    created.append((i, q))
        E = len(created)
        new_rect, new_cnt = [None] * E, [0] * E
        child_pos = {}
        n_expand = 0
        for c, (i, q) in enumerate(created):
            p = E - 1 - c
            r = rect[i]; hx = (r[2] - r[0] + 1) >> 1; hy = (r[3] - r[1] + 1) >> 1
            new_rect[p] = [(r[0], r[1], r[0] + hx, r[1] + hy), (r[0] + hx, r[1], r[2], r[1] + hy),
                           (r[0], r[1] + hy, r[0] + hx, r[3]), (r[0] + hx, r[1] + hy, r[2], r[3])][q]
            new_cnt[p] = cc[i][q]; child_pos[(i, q)] = p
            n_expand += cc[i][q] > 1
        kept_pos = {}
        for i in range(m):
            if i not in expd:
                kept_pos[i] = len(new_rect); new_rect.append(rect[i]); new_cnt.append(cnt[i])
        for k, (x, y, s) in enumerate(cands):
            nd = node_of[k]
            node_of[k] = child_pos[(nd, quadrant(x, y, rect[nd]))] if nd in expd else kept_pos[nd]
        prev, m = m, len(new_rect)
        rect, cnt, Efresh = new_rect, new_cnt, E
        if m >= N or m == prev:
            break
        if not phase2 and m + 3 * n_expand > N:
            phase2 = True
    best = [None] * m
    for k, (x, y, s) in enumerate(cands):
        nd = node_of[k]
        if best[nd] is None or s > cands[best[nd]][2]:
            best[nd] = k
    return best


if __name__ == "__main__":
    from oracle import oracle as orc
    import synth
    orc.build()
    bad = 0; total = 0
    rng = np.random.default_rng(0)
    # real candidate sets
    tex = synth.make_texture(0, 480, 640)
    for fi in range(6):
        img = synth.make_frame(tex, fi)
        for l, L in enumerate(orc.pyramid(img)):
            c = orc.fast_cells(L)
            h, w = L.shape
            N = int(orc.tables()["nfeat"][l])
            for NN in (N, max(5, N // 3), N * 4):
                ref = list(orc.distribute(c, 16, w - 16, 16, h - 16, NN))
                got = distribute_rounds([(int(a["x"]), int(a["y"]), int(a["score"])) for a in c], w - 32, h - 32, NN)
                total += 1; bad += ref != got
    # random sets, various aspect ratios
    for t in range(300):
        W = int(rng.integers(40, 1300)); H = int(rng.integers(40, 700))
        if round(W / H) < 1:
            continue
        n = int(rng.integers(1, 3000))
        pts = set()
        while len(pts) < min(n, W * H // 2):
            pts.add((int(rng.integers(0, W)), int(rng.integers(0, H))))
        c = np.zeros(len(pts), orc.CAND_DT)
        for i, (x, y) in enumerate(sorted(pts, key=lambda p: (p[1] // 30, p[0] // 30, p[1], p[0]))):
            c[i] = (x, y, int(rng.integers(7, 60)))
        N = int(rng.integers(1, 500))
        ref = list(orc.distribute(c, 16, W + 16, 16, H + 16, N))
        got = distribute_rounds([(int(a["x"]), int(a["y"]), int(a["score"])) for a in c], W, H, N)
        total += 1; bad += ref != got
    print("cases", total, "mismatches", bad)
