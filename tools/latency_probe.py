#!/usr/bin/env python3
"""One-frame-at-a-time drop-in calls (the bench's latency arm) with wall-clock per call; run under
`ncu --metrics gpu__time_duration.sum` to get the device time of every kernel of a single-frame chain."""
import sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "tests")]
import bench
ob = bench.load_pkg()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
frames, depths = bench.make_inputs(n, 0)
print(bench.latency_arm(ob, 0, frames, depths, n))
