#!/usr/bin/env python3
"""Quirk Q4 measured: how far the CPU path moves when it is compiled with the reference's own flags (-O3 -march=native: FMA contraction,
oracle/liborb_oracle_speed.so) instead of the parity flags (-O2 -ffp-contract=off, oracle/liborb_oracle.so).  CPU only; prints one JSON
object.  The parity build is the definition the CUDA path is held to bit for bit; this script says how much a reference binary itself
depends on its build."""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "tests")]
import synth  # noqa: E402
from oracle import oracle as orc  # noqa: E402

orc.build(); orc.build(speed=True)
tex = synth.make_texture(0, 480, 640)
kp = same_xy = desc_diff = bits = 0; ang = 0.0
for i in range(12):
    img = synth.make_frame(tex, i)
    k0, d0 = orc.extract(img); k1, d1 = orc.extract(img, speed=True)
    assert len(k0) == len(k1)
    kp += len(k0); same_xy += int(np.array_equal(k0["x"], k1["x"]) and np.array_equal(k0["y"], k1["y"]))
    da = np.abs(k0["angle"] - k1["angle"]); ang = max(ang, float(np.minimum(da, 360 - da).max()))
    nb = np.unpackbits(d0 ^ d1, axis=1).sum(1); desc_diff += int((nb > 0).sum()); bits += int(nb.sum())
pairs = []
for seed, outl in ((42, 0.3), (7, 0.6), (21, 0.6), (24, 0.3), (1234, 0.05), (5, 0.45), (3, 0.2), (11, 0.5)):
    src, dst, m, _, _ = synth.rigid_pairs(seed=seed, outlier_frac=outl)
    a = orc.ransac_iterate(src, dst, m, seed=seed, depth_cov=1.6e-3); b = orc.ransac_iterate(src, dst, m, seed=seed, depth_cov=1.6e-3, speed=True)
    pairs.append({"seed": seed, "outliers": outl, "same_inlier_list": a["inliers"].tobytes() == b["inliers"].tobytes(), "iterations": [a["real_iters"], b["real_iters"]],
                  "max_abs_T12_diff": float(np.abs(a["T12"] - b["T12"]).max())})
print(json.dumps({"extraction": {"frames": 12, "keypoints": kp, "frames_with_identical_positions": same_xy, "max_angle_diff_deg": ang,
                                 "descriptors_differing": desc_diff, "bits_differing": bits}, "ransac": pairs}))
