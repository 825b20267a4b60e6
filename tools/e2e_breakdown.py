#!/usr/bin/env python
"""Where the end-to-end step goes (authoring aid, run on the GPU box): plain H2D rate of the step's gray planes, the pipelined
sequence call alone, and the result downloads."""
import importlib.util, sys, time
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench
ob = bench.load_pkg()
F = 512
frames, depths = bench.make_inputs(F, 0)
ctx = ob.Context(max_frames=F, max_pairs=F)
hg = torch.from_numpy(frames).pin_memory(); hd = torch.from_numpy(depths).pin_memory()
dg = torch.empty_like(hg, device="cuda")
s = torch.cuda.Stream()
def timeit(f, n=20):
    for _ in range(3): f()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
def h2d():
    with torch.cuda.stream(s): dg.copy_(hg, non_blocking=True)
ms = timeit(h2d); print(f"plain H2D of {hg.numel()/1e6:.1f} MB: {ms:.2f} ms = {hg.numel()/ms/1e6:.1f} GB/s")
hgn, hdn = hg.numpy(), hd.numpy()
def seq():
    ctx.track_sequence(hgn, hdn, 0.8, True, seed=42); ctx.synchronize()
print(f"track_sequence + synchronize: {timeit(seq):.2f} ms")
def seq_nodepth():
    ctx.track_sequence(hgn, None, 0.8, True, ransac=False); ctx.synchronize()
print(f"track_sequence without depth / RANSAC: {timeit(seq_nodepth):.2f} ms")
def dl():
    ctx.download_ransac_summary(F - 1); ctx.match_counts(F - 1); ctx.frame_counts(F)
print(f"downloads: {timeit(dl):.3f} ms")
def ext():
    ctx.extract_batch(hgn, hdn); ctx.synchronize()
print(f"extract_batch (host, pipelined) + synchronize: {timeit(ext):.2f} ms")
