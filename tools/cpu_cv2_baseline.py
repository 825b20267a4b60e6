#!/usr/bin/env python
"""Sanity anchor for the CPU baseline ("B-cv2"): the oracle port against OpenCV's own single-threaded ORB + BFMatcher
on the same synthetic frames, same machine.  OpenCV's ORB is not the reference's extractor (no per-cell FAST, no quadtree), so this
says nothing about parity — only that the port bench.py times as `cpu_baseline` is not a straw man.
    python tools/cpu_cv2_baseline.py [n_frames]"""
import json
import sys
import time
from pathlib import Path

import cv2
import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "tests")]
import bench                                    # noqa: E402  (synthetic inputs of the bench workload)
from oracle import oracle as orc                # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 48
    frames, depths = bench.make_inputs(n, 0)
    cv2.setNumThreads(1)
    orb = cv2.ORB_create(nfeatures=1000, scaleFactor=1.2, nlevels=8, edgeThreshold=19, fastThreshold=20)
    bf = cv2.BFMatcher(cv2.NORM_HAMMING)
    t0 = time.perf_counter()
    prev = None; nk = 0; nm = 0
    for i in range(n):
        k, d = orb.detectAndCompute(frames[i], None)
        nk += len(k)
        if prev is not None and d is not None and len(d) >= 2:
            nm += sum(1 for m in bf.knnMatch(prev, d, k=2) if len(m) == 2 and m[0].distance < 0.8 * m[1].distance)
        prev = d
    t_cv = time.perf_counter() - t0
    orc.build()
    t0 = time.perf_counter()
    prev = None; nk2 = 0; nm2 = 0
    for i in range(n):
        k, d = orc.extract(frames[i], speed=True)
        nk2 += len(k)
        if prev is not None:
            nm2 += len(orc.knn_match(prev, d, 0.8, False, speed=True))
        prev = d
    t_or = time.perf_counter() - t0
    print(json.dumps({"frames": n, "workload": "extract + kNN-2 ratio match of consecutive frames, 640x480, 1000 kp, one thread",
                      "cv2_orb_bfmatcher": {"frames_per_s": n / t_cv, "mean_keypoints": nk / n, "mean_ratio_matches": nm / max(n - 1, 1), "opencv": cv2.__version__},
                      "oracle_port": {"frames_per_s": n / t_or, "mean_keypoints": nk2 / n, "mean_ratio_matches": nm2 / max(n - 1, 1)}}))


if __name__ == "__main__":
    main()
