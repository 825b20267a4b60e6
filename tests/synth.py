"""Seeded synthetic inputs for parity tests and bench.py.  numpy only.

Scene model: a textured fronto-parallel plane at depth z0 seen by a camera that translates parallel to
the plane and rolls about its optical axis, so consecutive frames are related by an exact 3D rigid motion
consistent with the (constant + noise + holes) depth map.  Intrinsics: TUM FR1 with distortion zeroed
(reference Utils/common.h:35-38,67; quirk Q11).
"""
import numpy as np

FX, FY, CX, CY = 517.3, 516.5, 318.6, 255.3
DEPTH_UNITS = 5000.0


def _gauss_kernel(sigma):
    r = int(3 * sigma + 0.5)
    x = np.arange(-r, r + 1, dtype=np.float64)
    k = np.exp(-0.5 * (x / sigma) ** 2)
    return (k / k.sum()).astype(np.float64)


def make_texture(seed, h, w, sigma=2.0, pad=128):
    rng = np.random.default_rng(seed)
    u = rng.integers(0, 256, size=(h + 2 * pad, w + 2 * pad)).astype(np.float64)
    k = _gauss_kernel(sigma)
    u = np.apply_along_axis(lambda r: np.convolve(r, k, mode="same"), 1, u)
    u = np.apply_along_axis(lambda c: np.convolve(c, k, mode="same"), 0, u)
    lo, hi = u.min(), u.max()
    return np.clip(np.rint((u - lo) * (255.0 / (hi - lo))), 0, 255).astype(np.uint8)


def motion(i, seed=0):
    """Smooth seeded motion: (tx, ty) pixels of texture shift and roll angle (rad) for frame i."""
    rng = np.random.default_rng(seed + 7919)
    ph = rng.uniform(0, 2 * np.pi, 3)
    tx = 40.0 * np.sin(0.11 * i + ph[0]) + 12.0 * np.sin(0.31 * i)
    ty = 30.0 * np.cos(0.07 * i + ph[1])
    th = np.deg2rad(3.0) * np.sin(0.05 * i + ph[2])
    return tx, ty, th


def make_frame(tex, i, w=640, h=480, seed=0, pad=128, noise_sigma=2.0, low_contrast_every=8):
    """Gray frame i: bilinear sample of the texture under motion(i) + N(0, sigma) noise."""
    tx, ty, th = motion(i, seed)
    ys, xs = np.mgrid[0:h, 0:w].astype(np.float64)
    # scale image coords about the principal point so that CX/CY stay meaningful for any w,h
    cx, cy = CX * w / 640.0, CY * h / 480.0
    c, s = np.cos(th), np.sin(th)
    u = c * (xs - cx) - s * (ys - cy) + cx + tx + pad
    v = s * (xs - cx) + c * (ys - cy) + cy + ty + pad
    u = np.clip(u, 0, tex.shape[1] - 2.001); v = np.clip(v, 0, tex.shape[0] - 2.001)
    u0 = np.floor(u).astype(np.int64); v0 = np.floor(v).astype(np.int64)
    fu = u - u0; fv = v - v0
    t = tex.astype(np.float64)
    val = (t[v0, u0] * (1 - fu) * (1 - fv) + t[v0, u0 + 1] * fu * (1 - fv)
           + t[v0 + 1, u0] * (1 - fu) * fv + t[v0 + 1, u0 + 1] * fu * fv)
    rng = np.random.default_rng(1000 + i + 100003 * seed)
    val = val + rng.normal(0.0, noise_sigma, size=val.shape)
    if low_contrast_every and i % low_contrast_every == low_contrast_every - 1:
        band = slice(h // 3, h // 3 + h // 4)
        val[band] = 128.0 + (val[band] - 128.0) * 0.35   # exercises the th=7 fallback cells
    return np.clip(np.rint(val), 0, 255).astype(np.uint8)


def make_depth(i, w=640, h=480, seed=0, z0=2.0, noise_m=0.002, hole_frac=0.03):
    rng = np.random.default_rng(5000 + i + 100003 * seed)
    z = z0 + rng.normal(0.0, noise_m, size=(h, w))
    d = np.clip(np.rint(z * DEPTH_UNITS), 1, 65535).astype(np.uint16)
    d[rng.random((h, w)) < hole_frac] = 0
    return d


def make_sequence(n, w=640, h=480, seed=0):
    tex = make_texture(seed, h, w)
    frames = np.stack([make_frame(tex, i, w, h, seed) for i in range(n)])
    depths = np.stack([make_depth(i, w, h, seed) for i in range(n)])
    return frames, depths


def descriptor_sets(n=1000, seed=1, match_frac=0.7, flip_p=0.08):
    """Config-2 descriptor sets: B holds noisy permuted copies of 70 % of A's rows + fresh rows."""
    rng = np.random.default_rng(seed)
    A = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    perm = rng.permutation(n)
    B = np.empty_like(A)
    nm = int(n * match_frac)
    bits = np.unpackbits(A[perm[:nm]], axis=1)
    flips = rng.random(bits.shape) < flip_p
    B[:nm] = np.packbits(bits ^ flips, axis=1)
    B[nm:] = np.random.default_rng(seed + 1).integers(0, 256, size=(n - nm, 32), dtype=np.uint8)
    order = rng.permutation(n)
    return A, B[order]


def tie_heavy_sets(n=1000, seed=3):
    """Bytes 2..31 zero: only 16 informative bits, so top-2 distance ties are everywhere (P5 order)."""
    rng = np.random.default_rng(seed)
    A = np.zeros((n, 32), np.uint8); B = np.zeros((n, 32), np.uint8)
    A[:, :2] = rng.integers(0, 256, size=(n, 2)); B[:, :2] = rng.integers(0, 256, size=(n, 2))
    return A, B


def rigid_pairs(m=650, seed=4, outlier_frac=0.3, n_pts=1000):
    """3D-3D correspondences for RANSAC: dst = R src + t + noise, a fraction replaced by outliers.
    Returns (src_xyz[n_pts,3], dst_xyz[n_pts,3], matches DMATCH[m], R, t)."""
    from oracle.oracle import DMATCH_DT
    rng = np.random.default_rng(seed)
    src = np.empty((n_pts, 3), np.float32)
    src[:, 2] = rng.uniform(0.8, 4.0, n_pts)
    src[:, 0] = rng.uniform(-0.6, 0.6, n_pts) * src[:, 2]
    src[:, 1] = rng.uniform(-0.45, 0.45, n_pts) * src[:, 2]
    ax = rng.normal(size=3); ax /= np.linalg.norm(ax)
    ang = np.deg2rad(4.0)
    K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
    R = np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * K @ K
    t = np.array([0.05, -0.03, 0.08])
    qi = rng.permutation(n_pts)[:m]
    ti = rng.permutation(n_pts)[:m]
    dst = np.zeros((n_pts, 3), np.float32)
    dst[:, 2] = rng.uniform(0.8, 4.0, n_pts)
    dst[:, 0] = rng.uniform(-0.6, 0.6, n_pts) * dst[:, 2]
    dst[:, 1] = rng.uniform(-0.45, 0.45, n_pts) * dst[:, 2]
    inl = rng.random(m) >= outlier_frac
    moved = (src[qi].astype(np.float64) @ R.T + t)
    moved += rng.normal(0, 0.002, size=moved.shape) * moved[:, 2:3]
    dst[ti[inl]] = moved[inl].astype(np.float32)
    # a few depth holes (z == 0) and NaNs to exercise the filter (ransac.cpp:175-189)
    holes = rng.permutation(m)[:10]
    src[qi[holes[:5]], :] = 0.0
    dst[ti[holes[5:8]], :] = 0.0
    dst[ti[holes[8:]], 2] = np.nan
    matches = np.zeros(m, DMATCH_DT)
    matches["queryIdx"] = qi; matches["trainIdx"] = ti
    matches["distance"] = rng.integers(5, 70, m).astype(np.float32)   # many ties, like Hamming
    return src, dst, matches, R, t
