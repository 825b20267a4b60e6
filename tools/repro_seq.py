#!/usr/bin/env python3
"""Debug aid: the bench sequence through track_sequence (pipelined / staged / device input) vs the oracle on the first NS frames."""
import sys, os
from pathlib import Path
from concurrent.futures import ThreadPoolExecutor
import numpy as np
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "tests")]
import bench
from oracle import oracle as orc
ob = bench.load_pkg(); orc.build()
NS = int(sys.argv[1]) if len(sys.argv) > 1 else 192
F = 512
frames, depths = bench.make_inputs(F, 0)
with ThreadPoolExecutor(os.cpu_count()) as ex:
    host = list(ex.map(lambda i: (lambda k, d: (k, d, orc.unproject(k, depths[i])[0]))(*orc.extract(frames[i])), range(NS)))
    matches = list(ex.map(lambda p: orc.knn_match(host[p][1], host[p + 1][1], 0.8, True), range(NS - 1)))
    r0 = orc.ransac_iterate(host[0][2], host[1][2], matches[0], seed=42)
    cov = r0["depth_cov"]
    rs = [r0] + list(ex.map(lambda p: orc.ransac_iterate(host[p][2], host[p + 1][2], matches[p], seed=42 + p, depth_cov=cov), range(1, NS - 1)))
hg = torch.from_numpy(frames).pin_memory(); hd = torch.from_numpy(depths.view(np.int16)).pin_memory()
dg = torch.from_numpy(frames).cuda(); dd = torch.from_numpy(depths.view(np.int16)).cuda()
for mode in ("pipelined", "pipelined", "staged", "device", "device"):
    ctx = ob.Context(max_frames=F, max_pairs=F, pipeline_chunk=-1 if mode == "staged" else 0)
    for rep in range(2):
        if mode == "device":
            ctx.track_sequence_device(dg.data_ptr(), 640, 640 * 480, F, dd.data_ptr(), 640, 640 * 480, 0.8, True, seed=42)
        else:
            ctx.track_sequence(hg.numpy(), hd.numpy().view(np.uint16), 0.8, True, seed=42)
        ctx.synchronize()
        bad = []
        for p in range(NS - 1):
            g = ctx.download_ransac(p); r = rs[p]
            if g["inliers"].tobytes() != r["inliers"].tobytes():
                bad.append((p, len(g["inliers"]), len(r["inliers"]), g["real_iters"], r["real_iters"], g["valid_iters"], r["valid_iters"], g["rmse"], r["rmse"], g["n_good"], r["n_good"]))
        print(mode, "rep", rep, "mismatching pairs:", bad)
    ctx.close()
