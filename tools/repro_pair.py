#!/usr/bin/env python3
"""Debug aid: one pair of the bench sequence through the device RANSAC (batched and standalone, with traces) against the oracle."""
import importlib.util, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "tests")]
import bench
from oracle import oracle as orc
ob = bench.load_pkg(); orc.build()
P = int(sys.argv[1]) if len(sys.argv) > 1 else 106
frames, depths = bench.make_inputs(512, 0)
ks = {}
for i in (0, 1, P, P + 1):
    k, d = orc.extract(frames[i]); ks[i] = (k, d, orc.unproject(k, depths[i])[0])
m0 = orc.knn_match(ks[0][1], ks[1][1], 0.8, True)
cov = orc.ransac_iterate(ks[0][2], ks[1][2], m0, seed=42)["depth_cov"]
m = orc.knn_match(ks[P][1], ks[P + 1][1], 0.8, True)
r = orc.ransac_iterate(ks[P][2], ks[P + 1][2], m, seed=42 + P, depth_cov=cov)
ctx = ob.Context(max_frames=4)
for want_table in (True, False):
    g = ctx.ransac_iterate(ks[P][2], ks[P + 1][2], m, seed=42 + P, depth_cov=cov, want_table=want_table)
    print("standalone want_table", want_table, "inliers equal", g["inliers"].tobytes() == r["inliers"].tobytes(), len(g["inliers"]), len(r["inliers"]),
          "iters", g["real_iters"], r["real_iters"], g["valid_iters"], r["valid_iters"], "rmse", g["rmse"], r["rmse"])
    hg, hr = g["hyp"], r["hyp"]
    for k in range(len(hr)):
        if hr[k]["rounds"] != hg[k]["rounds"] or hr[k]["n_refined"] != hg[k]["n_refined"] or hr[k]["refined_error"] != hg[k]["refined_error"]:
            print("  first differing hypothesis", k, "oracle", hr[k]["rounds"], hr[k]["n_refined"], hr[k]["refined_error"], "device", hg[k]["rounds"], hg[k]["n_refined"], hg[k]["refined_error"])
            break
    if want_table:
        print("  tables equal", np.array_equal(g["sample_table"], r["sample_table"]), "good equal", g["good_sorted"].tobytes() == r["good_sorted"].tobytes(), "M", r["n_good"])
# batched path on the 4 frames
fr = np.stack([frames[i] for i in (0, 1, P, P + 1)]); de = np.stack([depths[i] for i in (0, 1, P, P + 1)])
ctx.extract_batch(fr, de)
ctx.match_pairs(np.array([[0, 1], [2, 3]], np.int32), 0.8, True)
ctx.ransac_pairs(2, seed=42 + P - 1)          # pair slot 1 gets seed 42 + P
g = ctx.download_ransac(1)
print("batched inliers equal", g["inliers"].tobytes() == r["inliers"].tobytes(), len(g["inliers"]), len(r["inliers"]), g["real_iters"], r["real_iters"], g["depth_cov"], r["depth_cov"])
ctx.close()
