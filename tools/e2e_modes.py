#!/usr/bin/env python3
"""End-to-end step time (512 pinned host frames -> poses on the host) under the pipeline options: per-call (each call forks from and
joins the context stream, results read synchronously) vs overlapped (pipeline_overlap: two slot halves, results of step i read while
step i + 1 runs), chunk sizes, depth in place / staged / absent; next to the bare H2D time of the gray planes."""
import json, sys, time
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "tests")]
import bench
ob = bench.load_pkg()
F, W, H = 512, 640, 480
frames, depths = bench.make_inputs(F, 0)
hg = torch.from_numpy(frames).pin_memory(); hd = torch.from_numpy(depths.view(np.int16)).pin_memory()
g = hg.numpy(); d = hd.numpy().view(np.uint16)
out = {}
dst = torch.empty_like(hg, device="cuda")
torch.cuda.synchronize()
for _ in range(3): dst.copy_(hg, non_blocking=True)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(20): dst.copy_(hg, non_blocking=True)
torch.cuda.synchronize(); out["h2d_gray_only_ms"] = (time.perf_counter() - t0) / 20 * 1e3
STEPS = 30

def per_call(chunk, depth_mode):
    ctx = ob.Context(max_frames=F, max_pairs=F, pipeline_chunk=chunk, depth_zero_copy=-1 if depth_mode == "staged" else 0)
    dd = None if depth_mode == "none" else d
    rans = depth_mode not in ("none", "norans")
    def step():
        ctx.track_sequence(g, dd, 0.8, True, ransac=rans, seed=42)
        return ctx.download_ransac_summary(F - 1) if rans else ctx.match_counts(F - 1)
    for _ in range(3): step()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(STEPS): r = step()
    torch.cuda.synchronize(); ms = (time.perf_counter() - t0) / STEPS * 1e3
    ctx.close()
    return ms

def overlapped(chunk, depth_mode):
    ctx = ob.Context(max_frames=2 * F, max_pairs=2 * F, pipeline_chunk=chunk, pipeline_overlap=1, depth_zero_copy=-1 if depth_mode == "staged" else 0)
    dd = None if depth_mode == "none" else d
    rans = depth_mode not in ("none", "norans")
    fc = [torch.zeros(F, dtype=torch.int32).pin_memory() for _ in range(2)]
    mc = [torch.zeros(F - 1, dtype=torch.int32).pin_memory() for _ in range(2)]
    rr = [torch.zeros((F - 1) * ob.RANSAC_RESULT_DT.itemsize, dtype=torch.uint8).pin_memory() for _ in range(2)]
    def issue(i):
        h = i & 1
        ctx.track_sequence_at(g, dd, 0.8, h * F, h * F, True, ransac=rans, seed=42)
        ctx.read_results_async(h * F, h * F, F, fc[h].numpy(), mc[h].numpy(), rr[h].numpy().view(ob.RANSAC_RESULT_DT) if rans else None, h)
    for i in range(4):
        issue(i)
        if i: ctx.wait_marker((i - 1) & 1)
    ctx.wait_marker(1); torch.cuda.synchronize()
    t0 = time.perf_counter()
    issue(0)
    for i in range(1, STEPS):
        issue(i); ctx.wait_marker((i - 1) & 1)
    ctx.wait_marker((STEPS - 1) & 1)
    ms = (time.perf_counter() - t0) / STEPS * 1e3
    res = rr[(STEPS - 1) & 1].numpy().view(ob.RANSAC_RESULT_DT) if rans else None
    ok = float(res["ok"].mean()) if res is not None else None
    ctx.close()
    return ms, ok

for chunk in (64, 128, 256):
    for dm in ("inplace", "none", "norans"):
        out[f"per_call_chunk{chunk}_{dm}"] = per_call(chunk, dm)
        ms, ok = overlapped(chunk, dm)
        out[f"overlap_chunk{chunk}_{dm}"] = ms; out[f"overlap_chunk{chunk}_{dm}_ok"] = ok
print(json.dumps(out, indent=1))
