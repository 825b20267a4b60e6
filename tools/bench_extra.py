#!/usr/bin/env python3
"""Secondary measurements for the BASELINE configs that bench.py's headline line does not cover (one GPU, CUDA events /
wall clock around the C-ABI calls, inputs as stated).  Prints one JSON object; keep a copy under profiles/.

  config 1   extraction only, 640x480 / 1000 kp / 8 levels, frames in HBM
  config 4   1280x720: extraction with 2000 kp, and the adaptive-threshold FAST detector route (3x3 grid controllers)
  config 5   1000-keypoint query vs 2048 keyframes (many-to-many brute force on the device-resident store)
"""
import importlib.util
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "tests")]
import synth  # noqa: E402

spec = importlib.util.spec_from_file_location("orbfront_b200", ROOT / "adaptive-rgbd-localization-mappig_b200" / "__init__.py")
ob = importlib.util.module_from_spec(spec); sys.modules["orbfront_b200"] = ob; spec.loader.exec_module(ob)


def frames_of(w, h, n, seed=0, base=16):
    tex = synth.make_texture(seed, h, w)
    b = [synth.make_frame(tex, i, w, h, seed) for i in range(base)]
    return np.stack([b[i % base] for i in range(n)])


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


out = {}
# ---- config 1: extraction only -----------------------------------------------------------------------------------------
for (w, h, nf, n) in ((640, 480, 1000, 512), (1280, 720, 2000, 128)):
    fr = frames_of(w, h, n)
    d = torch.from_numpy(fr).cuda()
    ctx = ob.Context(width=w, height=h, nfeatures=nf, max_frames=n)
    def run():
        ctx.extract_batch_device(d.data_ptr(), w, w * h, n)
        ctx.synchronize()
    dt = timed(run)
    out[f"extract_{w}x{h}_{nf}kp"] = {"frames_per_s": n / dt, "ms_per_batch": dt * 1e3, "batch": n, "mean_keypoints": float(ctx.frame_counts(n).mean()),
                                      "input": "frames resident in HBM"}
    ctx.close()
# ---- config 4: adaptive-threshold FAST detector, 1280x720 ---------------------------------------------------------------
w, h, n = 1280, 720, 64
fr = frames_of(w, h, n)
ctx = ob.Context(width=w, height=h, nfeatures=2000, max_frames=1)
th = np.zeros(9)
t0 = time.perf_counter(); ctx.adaptive_detect(fr, th); cold = time.perf_counter() - t0        # thresholds start at 20 and climb
t0 = time.perf_counter(); kps, used, found = ctx.adaptive_detect(fr, th); dt = time.perf_counter() - t0   # controllers settled
out["adaptive_fast_1280x720"] = {"frames_per_s": n / dt, "ms_per_frame": dt / n * 1e3, "batch": n, "mean_keypoints": float(np.mean([len(k) for k in kps])),
                                 "cold_start_frames_per_s": n / cold, "final_thresholds": np.round(th, 2).tolist(), "input": "host frames (pageable), host keypoints out: H2D + D2H included"}
ctx.close()
# ---- config 5: query vs 2048 keyframes -----------------------------------------------------------------------------------
ctx = ob.Context(max_frames=8)
fr = frames_of(640, 480, 8)
ctx.extract_batch(fr)
nkf = 2048
ctx.kfdb_reserve(nkf)
for k in range(nkf):
    ctx.kfdb_add_from_slot(k, k % 8)
ctx.synchronize()
q = ctx.download_frame(3)[1]
counts = ctx.frame_counts(8)
pairs = float(len(q)) * float(sum(int(counts[k % 8]) for k in range(nkf)))
dt = timed(lambda: ctx.kfdb_match(q, 0, nkf, 0.8), reps=3, warm=1)
out["kfdb_query_vs_2048_keyframes"] = {"ms_per_query_incl_d2h_of_top2_tables": dt * 1e3, "descriptor_pairs": pairs, "pairs_per_s": pairs / dt,
                                       "note": "includes the D2H of 2048 x 1056 x 2 top-2 entries (17 MB) and host unpacking"}
ctx.close()
ctx2 = ob.Context(max_frames=8)
ctx2.extract_batch(fr)
ctx2.kfdb_reserve(nkf)
for k in range(nkf):
    ctx2.kfdb_add_from_slot(k, k % 8)
ctx2.synchronize()
dt = timed(lambda: ctx2.kfdb_survivors(q, 0, nkf, 0.8), reps=5, warm=2)
out["kfdb_query_vs_2048_keyframes_survivor_counts_only"] = {"ms_per_query": dt * 1e3, "descriptor_pairs": pairs, "pairs_per_s": pairs / dt,
                                                            "note": "H2D of the 32 KB query + kernels + D2H of 2048 survivor counts"}
ctx2.close()
print(json.dumps(out, indent=1))
