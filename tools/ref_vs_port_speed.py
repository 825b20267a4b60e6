#!/usr/bin/env python3
"""CPU anchor for bench.py's reference arm (authoring container only: needs oracle/_ref, i.e. the reference checkout): one core, the same
synthetic frames, three ways of running the reference path —
  ref     the reference's OWN orbextractor.cpp / matcher.cpp / ransac.cpp compiled verbatim (oracle/_ref, -O2 -ffp-contract=off) over
          the stand-in OpenCV / PCL / Eigen headers, whose numerical entry points are the oracle's routines;
  parity  the oracle port, parity build (-O2 -ffp-contract=off: same flags as ref);
  speed   the oracle port, -O3 -march=native (the reference's own flags, CMakeLists.txt:6) = what bench.py --impl reference times.
Says how much of the timed arm's speed is the port's own structure rather than the reference's (std::list quadtree, cv::Mat copies,
std::vector<std::vector<DMatch>> from knnMatch, ...)."""
import json
import sys
import time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "tests")]
import bench
from oracle import oracle as orc, ref

n = int(sys.argv[1]) if len(sys.argv) > 1 else 12
orc.build(); assert ref.available()
frames, depths = bench.make_inputs(n, 0)


def run_ref():
    prev = None; inl = 0
    for i in range(n):
        k, d = ref.extract(frames[i])
        xyz, _ = orc.unproject(k, depths[i])                       # Frame::ExtractFeatures' depth tail: a few microseconds, oracle in all three
        if prev is not None:
            m = ref.knn_match_frames(prev[0], d, bench.RATIO)
            r = ref.ransac_iterate(prev[1], xyz, m, seed=42 + i)
            inl += len(r["inliers"])
        prev = (d, xyz)
    return inl


def run_port(speed):
    prev = None; inl = 0; cov = -1.0
    for i in range(n):
        k, d = orc.extract(frames[i], speed=speed)
        xyz, _ = orc.unproject(k, depths[i])
        if prev is not None:
            m = orc.knn_match(prev[0], d, bench.RATIO, False, speed=speed)      # no cross-check: the reference's Matcher has none
            r = orc.ransac_iterate(prev[1], xyz, m, seed=42 + i, depth_cov=cov, speed=speed)
            cov = r["depth_cov"]; inl += len(r["inliers"])
        prev = (d, xyz)
    return inl


out = {"frames": n, "cores": 1, "what": "640x480, 1000 kp, extract + kNN-2 (ratio 0.8, no cross-check) + RANSAC(200,20,3.0,4), consecutive pairs"}
for name, fn in (("ref", run_ref), ("parity", lambda: run_port(False)), ("speed", lambda: run_port(True))):
    fn()
    t0 = time.perf_counter(); inl = fn(); dt = time.perf_counter() - t0
    out[name] = {"frames_per_s": n / dt, "ms_per_frame": 1e3 * dt / n, "inliers": int(inl)}
print(json.dumps(out))
