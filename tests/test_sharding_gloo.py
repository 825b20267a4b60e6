"""N > 1 host logic on CPU: world_size-2 gloo process groups run adaptive-rgbd-localization-mappig_b200/sharding.py with
the oracle standing in for the CUDA context (same method names and seed / depth-covariance semantics).  What is checked
is the sharding contract: every pair has exactly one owner, the halo frame makes straddling pairs computable, the depth
covariance latched by the globally first pair reaches every rank (quirk Q7), and the sharded run reproduces the
single-process sequence byte for byte.  Config 5: all-gather of keyframe shards and the gather-top-2 alternative."""
import importlib.util
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def load_sharding():
    spec = importlib.util.spec_from_file_location("orbf_sharding", ROOT / "adaptive-rgbd-localization-mappig_b200" / "sharding.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class OracleContext:
    """Context-shaped wrapper over the CPU oracle (tests only)."""

    def __init__(self):
        for p in (str(ROOT), str(ROOT / "tests")):
            if p not in sys.path:
                sys.path.insert(0, p)
        from oracle import oracle as orc
        orc.build()
        self.orc = orc
        self.latched = -1.0

    def extract_batch(self, frames, depths=None, slot0=0):
        self.frames = []
        for i in range(len(frames)):
            k, d = self.orc.extract(frames[i])
            xyz = self.orc.unproject(k, depths[i])[0] if depths is not None else None
            self.frames.append((k, d, xyz))

    def match_pairs(self, pairs, ratio, cross_check=False):
        self.pairs = [tuple(p) for p in np.asarray(pairs).reshape(-1, 2)]
        self.matches = [self.orc.knn_match(self.frames[a][1], self.frames[b][1], ratio, cross_check) for a, b in self.pairs]

    def ransac_pairs(self, npairs, seed=42, depth_cov=-1.0, **kw):
        self.results = []
        cov = depth_cov if depth_cov >= 0 else self.latched
        for k in range(npairs):
            a, b = self.pairs[k]
            r = self.orc.ransac_iterate(self.frames[a][2], self.frames[b][2], self.matches[k], seed=seed + k, depth_cov=cov)
            cov = r["depth_cov"]
            self.results.append(r)
        self.latched = cov

    def ransac_probe_depth_cov(self, npairs, seed=42, **kw):
        for k in range(npairs):
            a, b = self.pairs[k]
            cov = self.orc.ransac_iterate(self.frames[a][2], self.frames[b][2], self.matches[k], seed=seed + k, depth_cov=-1.0)["depth_cov"]
            if cov >= 0:
                return cov
        return -1.0

    def download_ransac(self, k):
        return dict(self.results[k])


def free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, n_frames, out_dir, dead_first=False):
    import torch
    import torch.distributed as dist
    sys.path[:0] = [str(ROOT), str(ROOT / "tests")]
    import synth
    sh_mod = load_sharding()
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        tex = synth.make_texture(0, 480, 640)
        sh = sh_mod.frame_shard(n_frames, world, rank)
        ids = range(sh["first"], sh["stop"])
        frames = np.stack([synth.make_frame(tex, i) for i in ids]); depths = np.stack([synth.make_depth(i) for i in ids])
        if dead_first and rank == 0:
            frames[0] = 128                     # rank 0's first pair has no matches: the covariance comes from a later pair
        ctx = OracleContext()
        shard, results, cov = sh_mod.run_sequence_shard(ctx, frames, depths, n_frames, rank, world, seed=42)
        np.savez(Path(out_dir) / f"rank{rank}.npz", cov=cov, pairs=np.array([r["pair"] for r in results]),
                 T=np.stack([r["T12"] for r in results]), ninl=np.array([len(r["inliers"]) for r in results]),
                 inl=np.concatenate([r["inliers"] for r in results]), rmse=np.array([r["rmse"] for r in results]))
        # trajectory: the composition rule chained rank to rank (each rank starts from the pose of its halo frame)
        T12 = np.stack([r["T12"] for r in results]) if results else np.zeros((0, 4, 4), np.float32)
        p0 = np.eye(4, dtype=np.float32); p0[:3, 3] = [0.5, -0.25, 0.125]
        poses = sh_mod.compose_trajectory_sharded(lambda npairs, start: ctx.orc.compose_trajectory(T12[:npairs], start), len(T12), rank, world, pose0=p0)
        np.save(Path(out_dir) / f"poses{rank}.npy", poses)
        # config 5: keyframe shards (2 keyframes per rank = this rank's first two frames), all-gather vs gather-top-2
        K = 1024
        local = torch.zeros((2, K, 32), dtype=torch.uint8); counts = torch.zeros(2, dtype=torch.int32)
        for j in range(2):
            d = ctx.frames[j][1]; local[j, :len(d)] = torch.from_numpy(d); counts[j] = len(d)
        gd, gc = sh_mod.gather_keyframes(local, counts)
        q = ctx.orc.extract(synth.make_frame(tex, 100))[1]

        def match_local(query, kf0, nkf, ratio):
            o = [np.full((nkf, len(query)), -1, np.int32) for _ in range(4)]; surv = np.zeros(nkf, np.int32)
            for j in range(nkf):
                t = local[kf0 + j, :int(counts[kf0 + j])].numpy()
                r = ctx.orc.knn2(query, t)
                for a in range(4):
                    o[a][j] = r[a]
                surv[j] = len(ctx.orc.knn_match(query, t, ratio))
            return o[0], o[1], o[2], o[3], surv
        i1, d1, i2, d2, surv = sh_mod.match_sharded_keyframes(match_local, q, 2, 0.8)
        ref1 = np.stack([ctx.orc.knn2(q, gd[j, :int(gc[j])].numpy())[0] for j in range(gd.shape[0])])
        assert gd.shape[0] == 2 * world and np.array_equal(i1, ref1), "gathered descriptors and gathered top-2 disagree"
        np.savez(Path(out_dir) / f"kf{rank}.npz", gd_sum=int(gd.to(torch.int64).sum()), gc=gc.numpy(), surv=surv)
    finally:
        dist.destroy_process_group()


def test_frame_shard_partitions_every_pair_once():
    sh = load_sharding()
    for n in (0, 1, 2, 5, 8, 4096, 4097):
        for world in (1, 2, 3, 8):
            owned = []
            frames = []
            for r in range(world):
                s = sh.frame_shard(n, world, r)
                assert 0 <= s["start"] <= s["stop"] <= n and s["first"] == s["start"] - s["halo"]
                frames += list(range(s["start"], s["stop"]))
                owned += list(range(*s["pairs"]))
                assert all(s["first"] <= p and p + 1 < s["stop"] for p in range(*s["pairs"])), "a pair's frames must be local"
            assert frames == list(range(n)) and owned == list(range(max(n - 1, 0))), (n, world)
    with pytest.raises(ValueError):
        sh.frame_shard(4, 2, 2)


@pytest.mark.parametrize("n,dead_first", [(5, False), (4, True)])
def test_two_rank_gloo_run_equals_single_process(tmp_path, orc, texture, n, dead_first):
    import torch.multiprocessing as mp
    import synth
    port = free_port()
    mp.spawn(_worker, args=(2, port, n, str(tmp_path), dead_first), nprocs=2, join=True)
    # single-process reference: the whole sequence on one rank
    ctx = OracleContext()
    frames = np.stack([synth.make_frame(texture, i) for i in range(n)]); depths = np.stack([synth.make_depth(i) for i in range(n)])
    if dead_first:
        frames[0] = 128
    _, ref, cov = load_sharding().run_sequence_shard(ctx, frames, depths, n, 0, 1, seed=42)
    got = [np.load(tmp_path / f"rank{r}.npz") for r in range(2)]
    assert all(float(g["cov"]) == cov for g in got), "depth covariance must be the globally first pair's on every rank"
    pairs = np.concatenate([g["pairs"] for g in got])
    assert list(pairs) == list(range(n - 1))
    T = np.concatenate([g["T"] for g in got]); ninl = np.concatenate([g["ninl"] for g in got]); rmse = np.concatenate([g["rmse"] for g in got])
    inl = np.concatenate([g["inl"] for g in got])
    # sharded trajectory == the single-process composition, bit for bit (rank 1's first pose is the halo frame = rank 0's last)
    p0 = np.eye(4, dtype=np.float32); p0[:3, 3] = [0.5, -0.25, 0.125]
    ref_poses = ctx.orc.compose_trajectory(np.stack([r["T12"] for r in ref]), p0)
    pr = [np.load(tmp_path / f"poses{r}.npy") for r in range(2)]
    assert np.array_equal(pr[1][0], pr[0][-1])
    assert np.array_equal(np.concatenate([pr[0], pr[1][1:]]), ref_poses)
    assert np.array_equal(T, np.stack([r["T12"] for r in ref])) and list(ninl) == [len(r["inliers"]) for r in ref]
    assert inl.tobytes() == np.concatenate([r["inliers"] for r in ref]).tobytes() and list(rmse) == [r["rmse"] for r in ref]
    kf = [np.load(tmp_path / f"kf{r}.npz") for r in range(2)]
    assert int(kf[0]["gd_sum"]) == int(kf[1]["gd_sum"]) and np.array_equal(kf[0]["gc"], kf[1]["gc"]) and np.array_equal(kf[0]["surv"], kf[1]["surv"])


def test_c_abi_frame_shard_equals_the_python_host_logic(ob):
    """orbf_frame_shard (what a C++ host calls) against sharding.frame_shard for every (n, world, rank) in a grid, incl. fewer frames than ranks."""
    import ctypes as C
    import importlib.util
    from pathlib import Path
    spec = importlib.util.spec_from_file_location("orbf_sharding", Path(__file__).resolve().parent.parent / "adaptive-rgbd-localization-mappig_b200" / "sharding.py")
    sh = importlib.util.module_from_spec(spec); spec.loader.exec_module(sh)
    L = ob.lib()
    out = [C.c_int32() for _ in range(6)]
    for n in (0, 1, 2, 3, 7, 8, 9, 64, 511, 4096):
        for world in (1, 2, 3, 4, 8):
            owners = []
            for rank in range(world):
                assert L.orbf_frame_shard(n, world, rank, *[C.byref(o) for o in out]) == 0
                want = sh.frame_shard(n, world, rank)
                got = dict(start=out[0].value, stop=out[1].value, halo=out[2].value, first=out[3].value, pairs=(out[4].value, out[5].value))
                assert got == want, (n, world, rank, got, want)
                owners += list(range(*got["pairs"]))
            assert owners == list(range(max(n - 1, 0))), (n, world)
    assert L.orbf_frame_shard(8, 0, 0, *[C.byref(o) for o in out]) != 0 and L.orbf_frame_shard(8, 2, 2, *[C.byref(o) for o in out]) != 0


def test_cpp_sequence_shard_host_logic(ob, tmp_path):
    """include/orbfront_shard.hpp (orbf::SequenceShard: the C++ statement of sharding.run_sequence_shard / compose_trajectory_sharded)
    over a recording stand-in of the device calls, ranks as threads (tests/cpp/shard_host_logic.cpp): for 0 .. 41 frames on 1 .. 8 ranks,
    with leading pairs that never score and with explicit covariances, the shards together give the results and absolute poses of one
    process bit for bit, every pair is solved once with seed + p and the covariance of the globally first scoring pair (quirks Q5 / Q7)."""
    import subprocess
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    pkg = root / "adaptive-rgbd-localization-mappig_b200"
    ob.lib()                                                     # orbf_frame_shard / orbf_status_string are the library's own
    exe = tmp_path / "shard_host_logic"
    r = subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", "-Werror", "-pthread", f"-I{root / 'include'}", "-o", str(exe),
                        str(root / "tests" / "cpp" / "shard_host_logic.cpp"), f"-L{pkg}", "-lorbfront_b200", f"-Wl,-rpath,{pkg}"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "360 sharded jobs identical to one process" in r.stdout, (r.stdout, r.stderr)
