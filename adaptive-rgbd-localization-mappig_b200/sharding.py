"""Multi-GPU host logic: one process per GPU (torch.distributed), frames partitioned with no data-path collective.

SURVEY.md §8(e): extraction is independent per frame, matching + RANSAC per consecutive frame pair.  A sequence of
n frames is cut into `world` contiguous chunks; rank r > 0 additionally extracts the last frame of rank r-1 (a halo
frame: 300 KB of H2D and one extra extraction, instead of a 44 KB peer copy on the critical path) so that the pair that
straddles two chunks has an owner.  The only scalar that crosses ranks is the depth covariance the reference latches in a
function-local static on the first pair it ever scores (quirk Q7, Odometry/ransac.cpp:416-421): every rank probes the value its
own pairs would latch (the first of them, in order, that reaches scoring — pair 0 of a shard may well have too few matches), the
8-byte candidates are all-gathered, and the first valid one in rank order is what every rank scores with.

BASELINE config 5 (keyframe-database many-to-many matching) is the one place with a real exchange step: every rank
holds a shard of the keyframe descriptors; `gather_keyframes` all-gathers the shards over NCCL (NVLink 5 / NVSwitch) into
one device buffer the matcher attaches to, `match_sharded_keyframes` keeps the shards in place and gathers only the per-
keyframe top-2 tables (16 KB per keyframe instead of 32 KB of descriptors).

Everything here takes plain tensors / numpy arrays and a process group, so the same code runs under gloo on CPU in
tests/test_sharding_gloo.py (world_size 2) with the oracle standing in for the CUDA context.
"""
import numpy as np


def frame_shard(n_frames, world, rank):
    """Contiguous chunk [start, stop) of rank `rank`, plus the halo frame it also extracts.

    Returns dict(start, stop, halo, first, pairs): `first` = first frame the rank extracts (start - halo), `pairs` =
    global pair indices [p0, p1) it owns (pair p = frames (p, p+1)).  Every pair 0..n_frames-2 has exactly one owner."""
    if world < 1 or not 0 <= rank < world or n_frames < 0:
        raise ValueError("bad shard request")
    base, extra = divmod(n_frames, world)
    start = rank * base + min(rank, extra)
    stop = start + base + (1 if rank < extra else 0)
    halo = 1 if (rank > 0 and stop > start and start > 0) else 0
    first = start - halo
    p0, p1 = (first, max(stop - 1, first)) if stop > start else (0, 0)
    return dict(start=start, stop=stop, halo=halo, first=first, pairs=(p0, p1))


def first_depth_cov(local_value, group=None, device="cpu"):
    """Quirk Q7 across ranks: every rank contributes the covariance its own pairs would latch (< 0 if none of them reaches scoring);
    the globally first pair that scores belongs to the lowest rank with a valid candidate."""
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(local_value)], dtype=torch.float64, device=device)
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1):
        return float(t.item())
    allv = torch.empty(dist.get_world_size(group), dtype=torch.float64, device=device)
    dist.all_gather_into_tensor(allv, t, group=group)
    valid = [float(v) for v in allv.cpu() if float(v) >= 0.0]
    return valid[0] if valid else -1.0


def run_sequence_shard(ctx, frames, depths, n_frames_total, rank, world, ratio=0.8, cross_check=True, seed=42, group=None,
                       device="cpu", first_pair_cov=None):
    """Extract + match + RANSAC this rank's shard of a sequence.  `frames` / `depths` hold exactly the frames
    [shard.first, shard.stop) in order.  `ctx` is a Context (CUDA) or anything with the same methods.

    Pair p of the global sequence is seeded with seed + p, exactly as one process running the whole sequence would."""
    sh = frame_shard(n_frames_total, world, rank)
    n_local = sh["stop"] - sh["first"]
    assert len(frames) == n_local, (len(frames), n_local)
    npairs = max(n_local - 1, 0)
    if n_local > 0:
        ctx.extract_batch(frames, depths)
    if npairs > 0:
        ctx.match_pairs(np.array([[i, i + 1] for i in range(npairs)], np.int32), ratio, cross_check)
    # depth covariance: the value latched by the globally first pair that reaches scoring, agreed on before any rank scores
    if first_pair_cov is not None:
        cov = first_pair_cov
    else:
        local = ctx.ransac_probe_depth_cov(npairs, seed=seed + sh["pairs"][0]) if npairs > 0 else -1.0
        cov = first_depth_cov(local, group, device)
    results = []
    if npairs > 0:
        # the library seeds pair slot k with seed + k: offset so that global pair p gets seed + p
        ctx.ransac_pairs(npairs, seed=seed + sh["pairs"][0], depth_cov=cov)
        for k in range(npairs):
            r = ctx.download_ransac(k)
            r["pair"] = sh["pairs"][0] + k
            results.append(r)
    return sh, results, cov


def compose_trajectory_sharded(compose_local, npairs_local, rank, world, pose0=None, group=None, device="cpu"):
    """Odometry::Compute's composition rule (pose[k+1] = T12[k] * pose[k], Odometry/odometry.cpp:82-84) over a sharded sequence.
    The float products are not associative, so a shard cannot be composed from the identity and corrected afterwards: rank r
    starts from the pose of its first frame — the halo frame, i.e. the last pose of rank r - 1 — which is passed down the ranks
    (64 bytes per hop, the only exchange of the sharded path besides the depth covariance).  compose_local(npairs, pose0) ->
    poses [npairs + 1, 4, 4] is Context.compose_trajectory (or the oracle's).  Returns this rank's poses, first = its first frame."""
    import torch
    import torch.distributed as dist
    multi = dist.is_available() and dist.is_initialized() and world > 1
    start = np.eye(4, dtype=np.float32) if pose0 is None else np.ascontiguousarray(pose0, np.float32).reshape(4, 4)
    if multi and rank > 0:
        t = torch.zeros(16, dtype=torch.float32, device=device)
        dist.recv(t, src=rank - 1, group=group)
        start = t.cpu().numpy().reshape(4, 4).copy()
    poses = np.asarray(compose_local(npairs_local, start), np.float32).reshape(-1, 4, 4)
    if multi and rank + 1 < world:
        dist.send(torch.from_numpy(np.ascontiguousarray(poses[-1]).reshape(16)).to(device), dst=rank + 1, group=group)
    return poses


class _CudaArrayView:
    """Zero-copy torch view of a raw device pointer (the library's keyframe store) through __cuda_array_interface__."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 3}


def device_tensor(ptr, shape, dtype="u1"):
    import torch
    return torch.as_tensor(_CudaArrayView(ptr, shape, {"u1": "|u1", "i4": "<i4"}[dtype]), device="cuda")


def gather_keyframes(local_desc, local_counts, group=None):
    """All-gather of the per-rank keyframe shards: local_desc [kf_local, K, 32] u8, local_counts [kf_local] i32 (torch
    tensors, CUDA under NCCL / CPU under gloo; every rank holds the same kf_local).  Returns (desc [world*kf_local, K, 32],
    counts [world*kf_local]) in rank order — the layout orbf_kfdb_attach_device expects."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local_desc, local_counts
    desc = torch.empty((world * local_desc.shape[0],) + tuple(local_desc.shape[1:]), dtype=local_desc.dtype, device=local_desc.device)
    counts = torch.empty(world * local_counts.shape[0], dtype=local_counts.dtype, device=local_counts.device)
    dist.all_gather_into_tensor(desc, local_desc.contiguous(), group=group)
    dist.all_gather_into_tensor(counts, local_counts.contiguous(), group=group)
    return desc, counts


def match_sharded_keyframes(match_local, query, kf_local, ratio, group=None, device="cpu"):
    """The alternative SURVEY.md §8(e) names: shards stay where they are, every rank matches the (replicated, 32 KB) query
    against its own keyframes, and only the per-keyframe top-2 tables + survivor counts are gathered.
    match_local(query, kf0, nkf, ratio) -> (idx1, d1, idx2, d2, survivors) as numpy arrays [nkf, nq] / [nkf]."""
    import torch
    import torch.distributed as dist
    i1, d1, i2, d2, surv = match_local(query, 0, kf_local, ratio)
    packed = torch.from_numpy(np.stack([i1, d1, i2, d2]).astype(np.int32)).to(device)          # [4, kf_local, nq]
    surv_t = torch.from_numpy(np.asarray(surv, np.int32)).to(device)
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return i1, d1, i2, d2, surv
    outs = [torch.empty_like(packed) for _ in range(world)]
    souts = [torch.empty_like(surv_t) for _ in range(world)]
    dist.all_gather(outs, packed, group=group)
    dist.all_gather(souts, surv_t, group=group)
    allp = torch.cat(outs, dim=1).cpu().numpy()
    return allp[0], allp[1], allp[2], allp[3], torch.cat(souts).cpu().numpy()
