#!/usr/bin/env python3
"""Per-stage device times (library CUDA events) of a 256-frame extraction + tracking step for the library named by ORBF_LIB (or the
default build): median of 7 profiled steps.  Prints one line: name, then stage = ms."""
import os, sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "tests")]
import bench
ob = bench.load_pkg()
F = int(os.environ.get("FRAMES", "256"))
frames, depths = bench.make_inputs(F, 0)
dg = torch.from_numpy(frames).cuda(); dd = torch.from_numpy(depths.view(np.int16)).cuda()
ctx = ob.Context(max_frames=F, max_pairs=F)
def step():
    ctx.track_sequence_device(dg.data_ptr(), 640, 640 * 480, F, dd.data_ptr(), 640, 640 * 480, 0.8, True, seed=42)
for _ in range(3): step()
ctx.synchronize()
ctx.profile_enable(True)
its = []; prev = {k: 0.0 for k in ctx.profile_read()}
for _ in range(7):
    step(); ctx.profile_collect()
    cur = {k: v[0] for k, v in ctx.profile_read().items()}
    its.append({k: cur[k] - prev[k] for k in cur}); prev = cur
ctx.profile_enable(False)
med = {k: float(np.median([i[k] for i in its])) for k in its[0]}
ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
st = torch.cuda.Stream(); ctx.set_stream(st.cuda_stream)
with torch.cuda.stream(st):
    step(); ev0.record(st)
    for _ in range(10): step()
    ev1.record(st)
torch.cuda.synchronize()
fc = ctx.frame_counts(F); mc = ctx.match_counts(F - 1)
print(os.path.basename(os.environ.get("ORBF_LIB", "default")), f"step={ev0.elapsed_time(ev1) / 10:.4f}", " ".join(f"{k}={v:.4f}" for k, v in med.items()), f"kp={int(fc.sum())} m={int(mc.sum())}")
ctx.close()
