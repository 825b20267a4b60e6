#!/usr/bin/env python3
"""Measurements for the SURVEY 8(f) "next" rows (one GPU): every entry point timed through the C ABI, host arrays in / host arrays out
(the call a reference-side binding makes, transfers and the call's synchronisation included; wall clock, median of `reps` calls after a
warm-up), with the CPU oracle on one host core beside it on the same inputs, and the results compared inside the run.  Prints one
JSON object; keep a copy under profiles/.

  8f-1  Matcher::ProjectionMatch / Fuse (search) / BoWMatch, Landmark::ComputeDistinctiveDescriptors
  8f-2  Frame::UndistortKeyPoints (cv::undistortPoints)
  8f-3  Odometry::Compute composition rule over a device-resident sequence
"""
import importlib.util
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "tests")]
import synth  # noqa: E402
from oracle import oracle as orc  # noqa: E402  (the checker beside the timed call; never inside it)
from test_fuse_bow import CAM, _bow_scene, _fuse_scene  # noqa: E402
from test_projection_match import _scene as _proj_scene  # noqa: E402

spec = importlib.util.spec_from_file_location("orbfront_b200", ROOT / "adaptive-rgbd-localization-mappig_b200" / "__init__.py")
ob = importlib.util.module_from_spec(spec); sys.modules["orbfront_b200"] = ob; spec.loader.exec_module(ob)
orc.build()


def med_ms(fn, reps=30, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    return float(np.median(ts) * 1e3)


def row(name, unit, units, gpu_fn, cpu_fn, same):
    g = med_ms(gpu_fn); c = med_ms(cpu_fn, reps=3, warm=1)
    return name, {"ms_per_call": round(g, 4), f"{unit}_per_s": units / (g * 1e-3), "cpu_oracle_ms_per_call": round(c, 3), "speedup_vs_one_core": round(c / g, 1),
                  "units_per_call": units, "identical_to_oracle": bool(same)}


out = {"what": "C-ABI calls, host arrays in / out, wall clock median; CPU oracle on one host core on the same inputs"}
ctx = ob.Context(max_frames=2)

# ---- ProjectionMatch: 1000 features, 2000 projected landmarks, radius 15 ----
sc = _proj_scene(12, n_feat=1000, n_lm=2000, radius=15.0)
kw = dict(radius=15.0, nn_ratio=0.8, th_high=100.0)
g = ctx.projection_match(sc[4], sc[5], sc[6], sc[7], kp_x=sc[0], kp_y=sc[1], kp_octave=sc[2], desc=sc[3], feat_taken=sc[8], **kw)
o = orc.projection_match(*sc[:8], feat_taken=sc[8], **kw)
k, v = row("projection_match_1000feat_2000lm", "landmarks", 2000,
           lambda: ctx.projection_match(sc[4], sc[5], sc[6], sc[7], kp_x=sc[0], kp_y=sc[1], kp_octave=sc[2], desc=sc[3], feat_taken=sc[8], **kw),
           lambda: orc.projection_match(*sc[:8], feat_taken=sc[8], **kw), np.array_equal(g[0], o[0]) and g[1] == o[1])
out[k] = v

# ---- Fuse search: 2000 features, 2000 landmarks ----
R, t, kx, ky, ur, desc, pw, lmd, valid = _fuse_scene(14, n_feat=2000, n_lm=2000, radius=5.0)
g = ctx.fuse_search(R, t, CAM, pw, lmd, valid, kp_x=kx, kp_y=ky, u_right=ur, desc=desc, radius=5.0, th_low=50)
o = orc.fuse_search(R, t, CAM, kx, ky, ur, desc, pw, lmd, valid, radius=5.0, th_low=50.0)
k, v = row("fuse_search_2000feat_2000lm", "landmarks", 2000,
           lambda: ctx.fuse_search(R, t, CAM, pw, lmd, valid, kp_x=kx, kp_y=ky, u_right=ur, desc=desc, radius=5.0, th_low=50),
           lambda: orc.fuse_search(R, t, CAM, kx, ky, ur, desc, pw, lmd, valid, radius=5.0, th_low=50.0), np.array_equal(g[0], o[0]) and np.array_equal(g[1], o[1]))
out[k] = v

# ---- BoWMatch: 1000 x 1000 features over 300 shared words ----
bs = _bow_scene(5, n1=1000, n2=1000, n_words=300)
g = ctx.bow_match(*bs); o = orc.bow_match(*bs)
k, v = row("bow_match_1000x1000_300words", "query_features", 1000, lambda: ctx.bow_match(*bs), lambda: orc.bow_match(*bs), g.tobytes() == o.tobytes())
out[k] = v

# ---- ComputeDistinctiveDescriptors: 4000 landmarks x 2..24 observations ----
rng = np.random.default_rng(3)
nobs = rng.integers(2, 25, 4000); offs = np.concatenate([[0], np.cumsum(nobs)]).astype(np.int32)
dd = rng.integers(0, 256, (int(offs[-1]), 32), dtype=np.uint8)
g = ctx.distinctive_descriptors(dd, offs); o = orc.distinctive_descriptors(dd, offs)
k, v = row("distinctive_descriptors_4000lm", "landmarks", 4000, lambda: ctx.distinctive_descriptors(dd, offs), lambda: orc.distinctive_descriptors(dd, offs),
           np.array_equal(g[0], o[0]))
out[k] = v

# ---- UndistortKeyPoints: 100 000 points, FR1 coefficients ----
xy = np.stack([rng.uniform(0, 640, 100000), rng.uniform(0, 480, 100000)], 1).astype(np.float32)
dist = np.array([0.262383, -0.953104, -0.005358, 0.002628, 1.163314], np.float32)
g = ctx.undistort_points(xy, 517.3, 516.5, 318.6, 255.3, dist); o = orc.undistort_points(xy, 517.3, 516.5, 318.6, 255.3, dist)
k, v = row("undistort_points_100k", "points", 100000, lambda: ctx.undistort_points(xy, 517.3, 516.5, 318.6, 255.3, dist),
           lambda: orc.undistort_points(xy, 517.3, 516.5, 318.6, 255.3, dist), np.array_equal(g, o))
out[k] = v
ctx.close()

# ---- Odometry::Compute composition rule + inlier flags over a tracked 64-frame sequence ----
n = 64
frames, depths = synth.make_sequence(n)
c2 = ob.Context(max_frames=n)
c2.track_sequence(frames, depths, 0.8, cross_check=True)
poses, flags = c2.compose_trajectory(n - 1)
T12 = np.stack([c2.download_ransac(p)["T12"] for p in range(n - 1)])
o = orc.compose_trajectory(T12)
k, v = row("compose_trajectory_63pairs", "poses", n - 1, lambda: c2.compose_trajectory(n - 1), lambda: orc.compose_trajectory(T12), np.array_equal(poses, o))
out[k] = v
c2.close()

print(json.dumps(out))
