import sys, numpy as np
sys.path.insert(0, '/root/repo')
import bench
ob = bench.load_pkg()
F = 512
frames, depths = bench.make_inputs(F, 0)
ctx = ob.Context(max_frames=F, max_pairs=F)
ctx.track_sequence(frames, depths, 0.8, True, seed=42); ctx.synchronize()
s = ctx.download_ransac_summary(F - 1)
ri = s["real_iters"]; vi = s["valid_iters"]
print("real_iters histogram:", np.bincount(ri)[:40].tolist())
print("valid_iters histogram:", np.bincount(vi)[:40].tolist())
print("max real_iters", ri.max(), "mean", ri.mean(), "n_inliers mean", s["n_inliers"].mean())
