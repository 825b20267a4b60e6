"""Second, independent oracle: the reference's extraction flow restated in Python on top of the SAME
OpenCV entry points the reference calls (cv2 4.13.0): cv2.resize(INTER_LINEAR), cv2.FastFeatureDetector per
cell ROI, cv2.GaussianBlur(7x7, sigma 2, REFLECT_101), cv2.fastAtan2, cv2.BFMatcher(NORM_HAMMING).knnMatch.
Used only by tests (to pin oracle/*.cpp) and by tools/make_goldens.py.

Follows /root/reference/Features/orbextractor.cpp: ComputePyramid :833-857, ComputeKeyPointsOctTree :665-746,
DistributeOctTree :466-663, IC_Angle :14-39, computeOrbDescriptor :43-85, operator() :756-815.
"""
import math

import cv2
import numpy as np

cv2.setNumThreads(1)
f32 = np.float32
EDGE = 19
HALF_PATCH = 15


def cv_round(v):
    return int(np.rint(v))


def scale_tables(nfeatures, scale_factor, nlevels):
    sf = float(f32(scale_factor))            # double member initialised from a float
    scale = [f32(1.0)]
    for _ in range(1, nlevels):
        scale.append(f32(float(scale[-1]) * sf))
    inv = [f32(1.0) / s for s in scale]
    factor = f32(1.0 / sf)
    desired = f32(nfeatures) * (f32(1) - factor) / (f32(1) - f32(math.pow(float(factor), float(nlevels))))
    nfeat, total = [], 0
    for _ in range(nlevels - 1):
        nfeat.append(cv_round(desired)); total += nfeat[-1]
        desired = f32(desired * factor)
    nfeat.append(max(nfeatures - total, 0))
    return scale, inv, nfeat


def pyramid(img, scale_factor=1.2, nlevels=8):
    _, inv, _ = scale_tables(1000, scale_factor, nlevels)
    h, w = img.shape
    levels = [img.copy()]
    for l in range(1, nlevels):
        sz = (cv_round(f32(w) * inv[l]), cv_round(f32(h) * inv[l]))
        levels.append(cv2.resize(levels[-1], sz, interpolation=cv2.INTER_LINEAR))
    return levels


def fast_cells(level, ini_th=20, min_th=7):
    """Per-cell cv::FAST with fallback; returns list of (x, y, response) relative to minBorder."""
    h, w = level.shape
    minB = EDGE - 3
    maxBX, maxBY = w - EDGE + 3, h - EDGE + 3
    width, height = f32(maxBX - minB), f32(maxBY - minB)
    nCols, nRows = int(width / f32(30)), int(height / f32(30))
    wCell, hCell = int(math.ceil(width / f32(nCols))), int(math.ceil(height / f32(nRows)))
    det_hi = cv2.FastFeatureDetector_create(ini_th, True)
    det_lo = cv2.FastFeatureDetector_create(min_th, True)
    out = []
    for i in range(nRows):
        iniY = minB + i * hCell
        maxY = iniY + hCell + 6
        if iniY >= maxBY - 3:
            continue
        maxY = min(maxY, maxBY)
        for j in range(nCols):
            iniX = minB + j * wCell
            maxX = iniX + wCell + 6
            if iniX >= maxBX - 6:
                continue
            maxX = min(maxX, maxBX)
            roi = level[iniY:maxY, iniX:maxX]
            kps = det_hi.detect(roi)
            if len(kps) == 0:
                kps = det_lo.detect(roi)
            for k in kps:
                out.append((int(k.pt[0]) + j * wCell, int(k.pt[1]) + i * hCell, int(k.response)))
    return out


class _Node:
    __slots__ = ("ulx", "uly", "urx", "bry", "keys", "no_more", "seq")


def _divide(n, cands):
    halfX = int(math.ceil(f32(n.urx - n.ulx) / f32(2)))
    halfY = int(math.ceil(f32(n.bry - n.uly) / f32(2)))
    ch = []
    for (ulx, uly, urx, bry) in ((n.ulx, n.uly, n.ulx + halfX, n.uly + halfY),
                                 (n.ulx + halfX, n.uly, n.urx, n.uly + halfY),
                                 (n.ulx, n.uly + halfY, n.ulx + halfX, n.bry),
                                 (n.ulx + halfX, n.uly + halfY, n.urx, n.bry)):
        c = _Node(); c.ulx, c.uly, c.urx, c.bry = ulx, uly, urx, bry; c.keys = []; c.no_more = False; c.seq = -1
        ch.append(c)
    mx, my = n.ulx + halfX, n.uly + halfY
    for k in n.keys:
        x, y = cands[k][0], cands[k][1]
        if x < mx:
            (ch[0] if y < my else ch[2]).keys.append(k)
        elif y < my:
            ch[1].keys.append(k)
        else:
            ch[3].keys.append(k)
    for c in ch:
        if len(c.keys) == 1:
            c.no_more = True
    return ch


def distribute(cands, minX, maxX, minY, maxY, N):
    """DistributeOctTree as a python list algorithm (list front = index 0).  Q3 tie rule: (count, seq)."""
    if not cands:
        return []
    nIni = int(math.floor(float(f32(maxX - minX) / f32(maxY - minY)) + 0.5))
    hX = f32(maxX - minX) / f32(nIni)
    seq = 0
    nodes = []
    for i in range(nIni):
        n = _Node()
        n.ulx, n.uly = int(hX * f32(i)), 0
        n.urx, n.bry = int(hX * f32(i + 1)), maxY - minY
        n.keys = []; n.no_more = False; n.seq = seq; seq += 1
        nodes.append(n)
    for k, c in enumerate(cands):
        nodes[int(f32(c[0]) / hX)].keys.append(k)
    nodes = [n for n in nodes if n.keys]
    for n in nodes:
        if len(n.keys) == 1:
            n.no_more = True
    finish = False
    while not finish:
        prev_size = len(nodes)
        n_expand = 0
        expandable = []
        front = []          # children pushed this pass; list order = reversed push order
        kept = []
        for n in nodes:
            if n.no_more:
                kept.append(n); continue
            for c in _divide(n, cands):
                if c.keys:
                    c.seq = seq; seq += 1
                    front.append(c)
                    if len(c.keys) > 1:
                        n_expand += 1; expandable.append(c)
        nodes = front[::-1] + kept
        if len(nodes) >= N or len(nodes) == prev_size:
            finish = True
        elif len(nodes) + n_expand * 3 > N:
            while not finish:
                prev_size = len(nodes)
                prev = sorted(expandable, key=lambda c: (len(c.keys), c.seq))
                expandable = []
                for n in reversed(prev):
                    pushed = []
                    for c in _divide(n, cands):
                        if c.keys:
                            c.seq = seq; seq += 1
                            pushed.append(c)
                            if len(c.keys) > 1:
                                expandable.append(c)
                    nodes.remove(n)
                    nodes = pushed[::-1] + nodes
                    if len(nodes) >= N:
                        break
                if len(nodes) >= N or len(nodes) == prev_size:
                    finish = True
    res = []
    for n in nodes:
        best = n.keys[0]
        for k in n.keys[1:]:
            if cands[k][2] > cands[best][2]:
                best = k
        res.append(best)
    return res


UMAX = [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]


def ic_angle(level, x, y):
    m01 = m10 = 0
    img = level.astype(np.int64)
    for u in range(-HALF_PATCH, HALF_PATCH + 1):
        m10 += u * int(img[y, x + u])
    for v in range(1, HALF_PATCH + 1):
        d = UMAX[v]
        us = np.arange(-d, d + 1)
        p = img[y + v, x - d:x + d + 1]; m = img[y - v, x - d:x + d + 1]
        m01 += v * int((p - m).sum())
        m10 += int((us * (p + m)).sum())
    return f32(cv2.fastAtan2(float(m01), float(m10)))


def rbrief(blurred, x, y, angle_deg, pattern):
    factor = f32(math.pi / 180.0)
    ang = f32(angle_deg) * factor
    a, b = f32(math.cos(float(ang))), f32(math.sin(float(ang)))
    pat = pattern.reshape(512, 2).astype(np.float32)
    px, py = pat[:, 0], pat[:, 1]
    # float32 products and sums, no FMA (numpy evaluates each op separately)
    yy = np.rint(px * b + py * a).astype(np.int64)
    xx = np.rint(px * a - py * b).astype(np.int64)
    vals = blurred[y + yy, x + xx].astype(np.int64).reshape(256, 2)
    bits = (vals[:, 0] < vals[:, 1]).astype(np.uint8).reshape(32, 8)
    return (bits << np.arange(8, dtype=np.uint8)).sum(axis=1).astype(np.uint8)


def extract(img, pattern, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7):
    scale, inv, nfeat = scale_tables(nfeatures, scale_factor, nlevels)
    levels = pyramid(img, scale_factor, nlevels)
    kps, descs, dbg = [], [], dict(levels=levels, cands=[], kept=[], blurred=[])
    for l, L in enumerate(levels):
        h, w = L.shape
        cands = fast_cells(L, ini_th, min_th)
        dbg["cands"].append(cands)
        keep = distribute(cands, EDGE - 3, w - EDGE + 3, EDGE - 3, h - EDGE + 3, nfeat[l])
        dbg["kept"].append(keep)
        if not keep:
            dbg["blurred"].append(None)
            continue
        blurred = cv2.GaussianBlur(L.copy(), (7, 7), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101)
        dbg["blurred"].append(blurred)
        size = f32(int(f32(31) * scale[l]))
        for idx in keep:
            x, y = cands[idx][0] + EDGE - 3, cands[idx][1] + EDGE - 3
            ang = ic_angle(L, x, y)
            descs.append(rbrief(blurred, x, y, ang, pattern))
            fx, fy = f32(x), f32(y)
            if l != 0:
                fx, fy = f32(fx * scale[l]), f32(fy * scale[l])
            kps.append((fx, fy, size, ang, f32(cands[idx][2]), l, -1))
    return kps, (np.stack(descs) if descs else np.zeros((0, 32), np.uint8)), dbg


def knn2(q, t):
    m = cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(q, t, k=2)
    i1 = np.array([r[0].trainIdx for r in m], np.int32); d1 = np.array([int(r[0].distance) for r in m], np.int32)
    i2 = np.array([r[1].trainIdx for r in m], np.int32); d2 = np.array([int(r[1].distance) for r in m], np.int32)
    return i1, d1, i2, d2
