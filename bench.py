#!/usr/bin/env python3
"""bench.py — frames/s of the ORB extract + Hamming match + RANSAC hot path on synthetic 640x480 RGB-D frames.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--frames F]            # this repo's CUDA path
  python bench.py --impl reference [--steps K] [--warmup W]                   # the CPU path on the host cores

A "step" is one pass of the hot path over one batch of F synthetic frames: pyramid, per-cell FAST, quadtree,
orientation, blur, rBRIEF, depth unprojection for every frame, then kNN-2 matching (ratio 0.8 + cross-check) and
RANSAC(200, 20, 3.0, 4) for the F-1 consecutive pairs (BASELINE.json configs[1] applied to a configs[2]-style
sequence shard; weak scaling: every rank processes its own F frames, no collective on the data path).
`value`   : inputs resident in HBM when the timed region starts, CUDA events on the launching stream, max over ranks.
`e2e`     : the same through the host-buffer C-ABI calls (pinned host frames -> H2D -> path -> D2H of the results).
`roofline`: dominant kernel, algorithmic bytes / CUDA-event time, against MEASURED_PEAKS.json (hbm_gbs).
`cpu_baseline`: the oracle ("port": the reference itself cannot be built here) on a bounded sample, 1 core.
"""
import argparse
import importlib.util
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "frames_per_sec_orb_extract_match_ransac_640x480_1000kp"
W, H = 640, 480
RATIO, CROSS = 0.8, True
# algorithmic HBM bytes per 640x480 frame and stage (u8, no border)
STAGE_BYTES = {"pyr_resize": 926_546 + 643_332, "fast_cell": 950_532, "blur7": 950_532 + 950_532}
FRAME_BYTES = 4_728_674


def config_object(F):
    """The workload both arms are measured on (identical object in both JSON lines; arm-specific detail lives in `arm`)."""
    return {"workload": "ORB extract (1000 kp, 8 levels, 1.2) + kNN-2 Hamming match (ratio 0.8, cross-check) + RANSAC(200,20,3.0,4), "
                        "640x480 RGB-D, consecutive pairs", "frames_per_step_per_gpu": F, "pairs_per_step_per_gpu": F - 1,
            "l2": f"no flush needed: per-step input {F * W * H * 3 / 1e6:.0f} MB > 126 MB L2",
            "sharding": "frames partitioned per rank, no data-path collective"}


def load_pkg():
    spec = importlib.util.spec_from_file_location("orbfront_b200", ROOT / "adaptive-rgbd-localization-mappig_b200" / "__init__.py")
    mod = importlib.util.module_from_spec(spec)
    sys.modules["orbfront_b200"] = mod
    spec.loader.exec_module(mod)
    return mod


def make_inputs(n_frames, seed):
    """n_frames synthetic frames: `base` distinct motion frames, reused cyclically with fresh sensor noise."""
    import synth
    tex = synth.make_texture(seed, H, W)
    base = min(n_frames, 64)
    frames = np.empty((n_frames, H, W), np.uint8); depths = np.empty((n_frames, H, W), np.uint16)
    clean = [synth.make_frame(tex, i, W, H, seed) for i in range(base)]
    dclean = [synth.make_depth(i, W, H, seed) for i in range(base)]
    rng = np.random.default_rng(seed + 999)
    for i in range(n_frames):
        j = i % (2 * base - 2) if base > 1 else 0
        j = j if j < base else 2 * base - 2 - j          # ping-pong so consecutive frames stay consecutive motions
        if i < base:
            frames[i] = clean[j]; depths[i] = dclean[j]
        else:
            noise = rng.integers(-2, 3, size=(H, W), dtype=np.int16)
            frames[i] = np.clip(clean[j].astype(np.int16) + noise, 0, 255).astype(np.uint8)
            depths[i] = dclean[j]
    return frames, depths


class ClockSampler:
    def __init__(self, gpu_index):
        self.rows = []; self.proc = None; self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu),
                 "--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
                 "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap",
                 "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True); self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, windows=None):
        """windows: [(t0, t1)] perf_counter intervals of the timed regions; samples inside them are the ones reported (a 100 ms
        sampling period against steps of a few ms: when no sample falls inside, the samples taken under load between the first
        warm-up step and the end of the last timed region are used and the line says so)."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = self.rows
        scope = "timed regions"
        if windows:
            inside = [r for r in rows if any(a <= r[0] <= b for a, b in windows)]
            if inside:
                rows = inside
            else:
                lo = getattr(self, "load_t0", None); hi = max(b for _, b in windows)
                rows = [r for r in rows if (lo is None or r[0] >= lo) and r[0] <= hi]
                scope = "under load (warm-up .. end of timed regions; no sample fell inside a timed region)"
        for _, r in rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[4 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "scope": scope}


def cpu_pipeline(orc, frames, depths, speed=True):
    """The reference's per-frame loop on the CPU oracle: extract, unproject, match with previous, RANSAC."""
    prev = None; cov = -1.0; n_inl = 0
    for i in range(len(frames)):
        k, d = orc.extract(frames[i], speed=speed)
        xyz, _ = orc.unproject(k, depths[i])
        if prev is not None:
            m = orc.knn_match(prev[0], d, RATIO, CROSS, speed=speed)
            r = orc.ransac_iterate(prev[1], xyz, m, seed=42 + i, depth_cov=cov, speed=speed)
            cov = r["depth_cov"]; n_inl += len(r["inliers"])
        prev = (d, xyz)
    return n_inl


def parity_in_bench(orc, ctx, frames, depths, ns, seed):
    """The timed path checked against the oracle inside the bench run: the first ns frames / ns - 1 pairs of the step that was just
    timed end to end (results still in the context) against the parity build of the oracle (no FMA contraction) on the same frames —
    keypoints, descriptors, 3D points, matches, inlier sets byte for byte, T12 within 1e-5 (the north star's tolerance)."""
    from concurrent.futures import ThreadPoolExecutor
    cores = os.cpu_count() or 1

    def ext(i):
        k, d = orc.extract(frames[i])
        return k, d, orc.unproject(k, depths[i])[0]
    with ThreadPoolExecutor(cores) as ex:
        host = list(ex.map(ext, range(ns)))
        matches = list(ex.map(lambda p: orc.knn_match(host[p][1], host[p + 1][1], RATIO, CROSS), range(ns - 1)))
        r0 = orc.ransac_iterate(host[0][2], host[1][2], matches[0], seed=seed, depth_cov=-1.0)
        cov = r0["depth_cov"]                   # quirk Q7: latched by the first pair that scores, used by every later one
        rs = [r0] + list(ex.map(lambda p: orc.ransac_iterate(host[p][2], host[p + 1][2], matches[p], seed=seed + p, depth_cov=cov), range(1, ns - 1)))
    bad = []
    for i in range(ns):
        k, d, xyz = ctx.download_frame(i)
        if k.tobytes() != host[i][0].tobytes() or not np.array_equal(d, host[i][1]) or not np.array_equal(xyz, host[i][2]):
            bad.append(f"frame {i}")
    max_dt = 0.0
    for p in range(ns - 1):
        g = ctx.download_ransac(p); r = rs[p]
        if ctx.download_matches(p).tobytes() != matches[p].tobytes():
            bad.append(f"pair {p}: matches")
        if g["ok"] != r["ok"] or g["inliers"].tobytes() != r["inliers"].tobytes():
            bad.append(f"pair {p}: inliers")
        dt = float(np.abs(g["T12"] - r["T12"]).max()); max_dt = max(max_dt, dt)
        if not dt <= 1e-5:
            bad.append(f"pair {p}: T12")
    return {"frames": ns, "pairs": ns - 1, "identical": not bad, "max_abs_T12_diff": max_dt, "mismatches": bad[:8],
            "checked": "keypoints, descriptors, 3D points, matches, RANSAC inlier sets (bytes); T12 <= 1e-5; oracle parity build, end-to-end step"}


def run_reference(args):
    """--impl reference: the CPU implementation of the path with all host threads (frame-chunk parallel)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from concurrent.futures import ThreadPoolExecutor
    from oracle import oracle as orc
    orc.build()
    cores = os.cpu_count() or 1
    per = 6                                     # frames per worker chunk: 6 extractions + 5 pair matches/RANSACs
    n = cores * per
    frames, depths = make_inputs(n, 0)
    chunks = [(frames[i * per:(i + 1) * per], depths[i * per:(i + 1) * per]) for i in range(cores)]

    def step():
        with ThreadPoolExecutor(cores) as ex:      # ctypes releases the GIL: real thread parallelism
            list(ex.map(lambda c: cpu_pipeline(orc, c[0], c[1]), chunks))

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    fps = n / dt
    sample = f"{n} synthetic 640x480 RGB-D frames per step in {cores} chunks of {per} (extract all, match+RANSAC {per - 1} pairs per chunk)"
    out = {"impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "u8", "data": "synthetic",
           "config": config_object(args.frames),
           "arm": {"frames_per_step": n, "note": "CPU oracle port of the reference path (bit-identical to the reference's own orbextractor.cpp / matcher.cpp / "
                                                   "ransac.cpp compiled verbatim in oracle/_ref over stand-in OpenCV / PCL / Eigen headers; the real libraries are "
                                                   "not in the image, so the timed arm is the port), g++ -O3 -march=native; "
                                                   "each step is a bounded sample of the workload"},
           "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
           "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


def run_config5(args):
    """--config 5: keyframe-database many-to-many matching (BASELINE configs[4]): every rank holds 2048 / N keyframes (~1000 rows of 32
    bytes each) and matches its own 1000-keypoint query against ALL 2048.  Two ways to reach the other ranks' keyframes, both timed with
    CUDA events on the context stream, max over ranks:  allgather = ncclAllGather of the stores into a gathered copy + local match;
    peers = no collective, the matcher reads every keyframe from its owner's HBM over NVLink (CUDA IPC).  Parity: a sample of
    keyframes from every rank against the oracle (top-2 tables and ratio survivors)."""
    import torch
    import torch.distributed as dist
    import synth
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        sys.stdout.flush(); saved = os.dup(1); os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local)); dist.barrier(); torch.cuda.synchronize()
        finally:
            sys.stdout.flush(); os.dup2(saved, 1); os.close(saved)
    ob = load_pkg()
    NKF, NBASE = 2048, 16
    kf_local = NKF // world
    tex = synth.make_texture(0, H, W)
    base = np.stack([synth.make_frame(tex, 2 * i, W, H, 0) for i in range(NBASE)])          # "every 2nd frame's output"
    qframe = synth.make_frame(tex, 2 * rank + 1, W, H, 0)[None]
    ctx = ob.Context(max_frames=NBASE + 1, device=local)
    stream = torch.cuda.Stream(device=local); ctx.set_stream(stream.cuda_stream)
    ctx.extract_batch(base); ctx.extract_batch(qframe, slot0=NBASE)
    counts = ctx.frame_counts(NBASE + 1)
    nq = int(counts[NBASE])
    ctx.kfdb_reserve(kf_local)
    for k in range(kf_local):
        ctx.kfdb_add_from_slot(k, (rank * kf_local + k) % NBASE)       # global keyframe g holds base frame g % NBASE
    ctx.synchronize()
    kf_rows = np.array([counts[g % NBASE] for g in range(NKF)], np.int64)
    pairs_per_query = float(nq * kf_rows.sum())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn):
        for _ in range(args.warmup):
            fn()
        barrier()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            for _ in range(args.steps):
                fn()
            e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1) / args.steps
        if world > 1:
            t = torch.tensor([ms], device=f"cuda:{local}", dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t[0])
        return ms

    out = {}
    results = {}
    if world > 1:
        idt = torch.from_numpy(ob.comm_unique_id() if rank == 0 else np.zeros(128, np.uint8)).cuda(local)
        dist.broadcast(idt, 0)
        ctx.comm_init(idt.cpu().numpy(), world, rank)
        assert ctx.kfdb_allgather() == NKF
        gather_ms = timed(lambda: ctx.kfdb_allgather())
        match_ms = timed(lambda: ctx.kfdb_match_slot(NBASE, 0, NKF, 0.8))
        both_ms = timed(lambda: (ctx.kfdb_allgather(), ctx.kfdb_match_slot(NBASE, 0, NKF, 0.8)))
        results["allgather"] = ctx.kfdb_results(NKF, nq)
        gbytes = kf_local * ctx.K * 32
        out["allgather"] = {"gather_ms": gather_ms, "match_ms": match_ms, "gather_plus_match_ms": both_ms, "bytes_per_rank": gbytes,
                            "bus_GBps": gbytes * (world - 1) / (gather_ms * 1e-3) / 1e9, "nvlink5_peak_GBps_per_direction": 900.0,
                            "frac_of_nvlink": gbytes * (world - 1) / (gather_ms * 1e-3) / 1e9 / 900.0,
                            "pairs_per_s": world * pairs_per_query / (both_ms * 1e-3)}
        a, b = ctx.kfdb_ipc_handles()
        mine = torch.from_numpy(np.concatenate([a, b])).cuda(local); allh = torch.empty((world, 128), dtype=torch.uint8, device=f"cuda:{local}")
        dist.all_gather_into_tensor(allh, mine)
        hh = allh.cpu().numpy()
        ctx.kfdb_attach_peers(hh[:, :64].copy(), hh[:, 64:].copy(), world, rank, kf_local)
        peers_ms = timed(lambda: ctx.kfdb_match_slot(NBASE, 0, NKF, 0.8))
        results["peers"] = ctx.kfdb_results(NKF, nq)
        out["peers"] = {"match_ms": peers_ms, "pairs_per_s": world * pairs_per_query / (peers_ms * 1e-3),
                        "remote_bytes_per_rank": int((world - 1) * kf_local * ctx.K * 32),
                        "remote_read_GBps": (world - 1) * kf_local * ctx.K * 32 / (peers_ms * 1e-3) / 1e9}
        headline_ms, how = min((both_ms, "allgather + match"), (peers_ms, "peer reads, no collective"))
    else:
        headline_ms = timed(lambda: ctx.kfdb_match_slot(NBASE, 0, NKF, 0.8))
        how = "single GPU holds all keyframes"
        results["local"] = ctx.kfdb_results(NKF, nq)
        out["local"] = {"match_ms": headline_ms, "pairs_per_s": pairs_per_query / (headline_ms * 1e-3)}
    # parity against the oracle: 3 keyframes of every rank's shard (global ids), every route measured
    from oracle import oracle as orc
    orc.build()
    bdesc = [orc.extract(base[i])[1] for i in range(NBASE)]
    q = orc.extract(qframe[0])[1]
    bad = []
    sample = sorted({r * kf_local + j for r in range(world) for j in (0, kf_local // 2, kf_local - 1)})
    for route, (i1, d1, i2, d2, surv) in results.items():
        for g in sample:
            ref = orc.knn2(q, bdesc[g % NBASE])
            if not (np.array_equal(i1[g], ref[0]) and np.array_equal(d1[g], ref[1]) and np.array_equal(i2[g], ref[2]) and np.array_equal(d2[g], ref[3])
                    and int(surv[g]) == len(orc.knn_match(q, bdesc[g % NBASE], 0.8))):
                bad.append(f"{route}: keyframe {g}")
    ok = torch.tensor([0 if bad else 1], device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if rank == 0:
        value = world * pairs_per_query / (headline_ms * 1e-3)
        print(json.dumps({"metric": "descriptor_pairs_per_sec_kfdb_1000kp_query_vs_2048_keyframes", "value": value, "unit": "pairs/s", "n_gpus": world,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": headline_ms, "higher_is_better": True, "scaling": "strong",
                          "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                          "config": {"workload": "BASELINE configs[4]: keyframe-database many-to-many Hamming kNN-2 + ratio 0.8, one ~1000-keypoint query per rank against 2048 "
                                                 "keyframes sharded over the ranks", "keyframes": NKF, "keyframes_per_rank": kf_local, "query_rows": nq,
                                     "mean_keyframe_rows": float(kf_rows.mean()), "route": how},
                          "routes": out, "parity": {"keyframes_checked_per_route": len(sample), "identical_on_every_rank": bool(int(ok[0])), "mismatches": bad[:6]},
                          "gpu_launches": int(ctx.launch_count())}))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    if bad:
        raise SystemExit("bench.py --config 5: device results disagree with the oracle: " + "; ".join(bad[:6]))


def run_config4(args):
    """--config 4 (BASELINE configs[3]): 1280x720 frames, 2000 keypoints, 8 levels, detector thresholds adapted per 3x3 image region by
    stateful controllers.  One GPU per rank, every rank its own clip (weak scaling, no collective).  Lines of the object:
      value / extract_adapted : the 8-level ORB extractor with iniThFAST driven by the controller state (orbf_extract_adapted, host frames in,
                                controller step chained per frame on the device) — frames/s, wall clock around the call + synchronize;
      extract_fixed_device    : the same extractor with the global thresholds on frames resident in HBM (CUDA events);
      chain_device            : extract + kNN-2 (ratio 0.8, cross-check) + RANSAC over consecutive pairs at 2000 keypoints (CUDA events);
      adaptive_detect         : the reference's own config-4 code path — single-scale FAST with the grid / dynamic adjusters
                                (videogridadaptedfeaturedetector.cpp:52-84), host frames in, host keypoints out.
    Parity inside the run: the first frames of the adapted clip against the oracle (keypoints, descriptors, thresholds used)."""
    import torch
    import torch.distributed as dist
    import synth
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        sys.stdout.flush(); saved = os.dup(1); os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local)); dist.barrier(); torch.cuda.synchronize()
        finally:
            sys.stdout.flush(); os.dup2(saved, 1); os.close(saved)
    ob = load_pkg()
    w, h, nf = 1280, 720, 2000
    F = min(args.frames, 128)
    tex = synth.make_texture(rank, h, w)
    base = [synth.make_frame(tex, i, w, h, rank) for i in range(16)]
    frames = np.stack([base[i % 30 if i % 30 < 16 else 30 - i % 30] for i in range(F)])             # ping-pong: consecutive motions
    depths = np.stack([synth.make_depth(i % 16, w, h, rank) for i in range(F)])
    band = dict(min_features=round(0.6 * nf / 9), max_features=round(1.02 * nf / 9))
    ctx = ob.Context(width=w, height=h, nfeatures=nf, max_frames=F, max_pairs=F, device=local)
    stream = torch.cuda.Stream(device=local); ctx.set_stream(stream.cuda_stream)
    steps = max(1, min(args.steps, 20))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def wall(fn):
        for _ in range(max(1, min(args.warmup, 3))):
            fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) * 1e3 / steps

    def events(fn):
        for _ in range(max(1, min(args.warmup, 3))):
            fn()
        barrier()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            for _ in range(steps):
                fn()
            e1.record(stream)
        barrier()
        return e0.elapsed_time(e1) / steps

    hfr = torch.from_numpy(frames).pin_memory().numpy()
    th = np.zeros(9)

    def adapted():
        ctx.extract_adapted(hfr, th, **band); ctx.synchronize()
    adapted_ms = wall(adapted)
    l0 = ctx.launch_count(); adapted(); adapted_launches = ctx.launch_count() - l0
    kp_adapted = float(ctx.frame_counts(F).mean())
    # the same chain with V independent videos advancing together (V frames per link): what a multi-camera rig / several sequences give
    V = 16 if F % 16 == 0 else 1
    T = F // V
    vids = hfr.reshape(V, T, h, w)
    thv = np.zeros((V, 9))

    def adapted_videos():
        ctx.extract_adapted_videos(vids, thv, **band); ctx.synchronize()
    videos_ms = wall(adapted_videos)
    d_gray = torch.from_numpy(frames).cuda(local); d_depth = torch.from_numpy(depths.view(np.int16)).cuda(local)
    fixed_ms = events(lambda: ctx.extract_batch_device(d_gray.data_ptr(), w, w * h, F))
    kp_fixed = float(ctx.frame_counts(F).mean())
    chain_ms = events(lambda: ctx.track_sequence_device(d_gray.data_ptr(), w, w * h, F, d_depth.data_ptr(), w, w * h, RATIO, CROSS, seed=42))
    summ = ctx.download_ransac_summary(F - 1); mc = ctx.match_counts(F - 1)
    nd = min(F, 64)
    dctx = ob.Context(width=w, height=h, nfeatures=nf, max_frames=1, device=local)
    th2 = np.zeros(9)
    dctx.adaptive_detect(frames[:nd], th2)                  # thresholds settle
    t0 = time.perf_counter(); kps, used, found = dctx.adaptive_detect(frames[:nd], th2); det_ms = (time.perf_counter() - t0) * 1e3
    dctx.close()
    # parity: a fresh controller state over the first frames against the oracle
    parity = None
    if rank == 0 and args.cpu_sample > 0:
        from oracle import oracle as orc
        orc.build()
        ns = min(4, F)
        cfg = orc.adaptive_default(); cfg.min_features, cfg.max_features = band["min_features"], band["max_features"]
        th_ref = np.zeros(9); th_got = np.zeros(9)
        t0 = time.perf_counter()
        ref = [orc.extract_adapted(f, th_ref, nfeatures=nf, cfg=cfg) for f in frames[:ns]]
        cpu_dt = time.perf_counter() - t0
        used_g, found_g = ctx.extract_adapted(frames[:ns], th_got, **band)
        bad = []
        for i in range(ns):
            k, d, _ = ctx.download_frame(i)
            if k.tobytes() != ref[i][0].tobytes() or not np.array_equal(d, ref[i][1]) or not np.array_equal(used_g[i], ref[i][2]):
                bad.append(f"frame {i}")
        if not np.array_equal(th_ref, th_got):
            bad.append("controller state")
        parity = {"frames": ns, "identical": not bad, "mismatches": bad, "checked": "keypoints, descriptors, thresholds used per region, controller state"}
        cpu = {"value": ns / cpu_dt, "unit": "frames/s", "cores": 1, "kind": "port", "sample": f"oracle extract_adapted on the first {ns} frames, {cpu_dt:.1f} s"}
    else:
        cpu = None
    vals = torch.tensor([adapted_ms, fixed_ms, chain_ms, det_ms / nd, videos_ms], device=f"cuda:{local}", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
    adapted_ms, fixed_ms, chain_ms, det_ms_per, videos_ms = [float(x) for x in vals]
    if rank == 0:
        px = sum(int(round(w / 1.2 ** l)) * int(round(h / 1.2 ** l)) for l in range(8))
        out = {"metric": "frames_per_sec_orb_extract_1280x720_2000kp_adapted_thresholds", "value": world * F / (adapted_ms * 1e-3), "unit": "frames/s",
               "n_gpus": world, "steps": steps, "warmup": args.warmup, "ms_per_step": adapted_ms, "higher_is_better": True, "scaling": "weak",
               "vs_baseline": None, "dtype": "u8", "data": "synthetic",
               "config": {"workload": "BASELINE configs[3]: 1280x720, 2000 kp, 8 levels, scale 1.2, iniThFAST adapted per 3x3 region by stateful controllers "
                                      "(x0.7 below / x1.3 above the band), frames in video order", "frames_per_step_per_gpu": F,
                          "band": band, "l2": f"per-step input {F * w * h / 1e6:.0f} MB of gray planes from pinned host memory"},
               "routes": {"extract_adapted": {"frames_per_s": world * F / (adapted_ms * 1e-3), "ms_per_step": adapted_ms, "mean_keypoints": kp_adapted,
                                              "input": "pinned host frames, H2D inside the timed region; per-frame controller step on the device",
                                              "launches_per_step": int(adapted_launches)},
                          "extract_adapted_videos": {"frames_per_s": world * F / (videos_ms * 1e-3), "ms_per_step": videos_ms, "videos": V, "frames_per_video": T,
                                                     "input": "pinned host frames of V independent videos, H2D inside the timed region; V frames per link of the "
                                                              "FAST -> quadtree -> controller chain (orbf_extract_adapted_videos)"},
                          "extract_fixed_device": {"frames_per_s": world * F / (fixed_ms * 1e-3), "ms_per_step": fixed_ms, "mean_keypoints": kp_fixed,
                                                   "input": "frames resident in HBM", "algorithmic_bytes_per_frame": 14_193_683,
                                                   "frac_of_hbm": 14_193_683 * F / (fixed_ms * 1e-3) / 1e9 / 6538.9},
                          "chain_device": {"frames_per_s": world * F / (chain_ms * 1e-3), "ms_per_step": chain_ms, "mean_matches": float(mc.mean()),
                                           "ransac_ok_frac": float(np.mean(summ["ok"])), "mean_inliers": float(np.mean(summ["n_inliers"])),
                                           "what": "extract + kNN-2 (ratio 0.8, cross-check) + RANSAC(200,20,3.0,4), consecutive pairs, inputs in HBM"},
                          "adaptive_detect": {"frames_per_s": world * 1e3 / det_ms_per, "ms_per_frame": det_ms_per, "mean_keypoints": float(np.mean([len(k) for k in kps])),
                                              "what": "single-scale FAST + grid / dynamic adjusters (the reference's own config-4 route), host frames in, "
                                                      "host keypoints out, controllers settled"}},
               "pixels_per_frame_all_levels": px, "gpu_launches": int(adapted_launches), "cpu_baseline": cpu, "parity_in_bench": parity}
        print(json.dumps(out))
        if parity is not None and not parity["identical"]:
            raise SystemExit("bench.py --config 4: device results disagree with the oracle: " + "; ".join(parity["mismatches"]))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def latency_arm(ob, local, frames, depths, n):
    """The reference's per-frame tracking loop (System/tracking.cpp:38-46,193-208; protocol of Tests/detector-descriptor-speed-test.cpp:53-70)
    through the drop-in calls, ONE frame at a time from pageable host memory: extraction + depth unprojection of the frame with every
    result copied back to host vectors, Matcher::KnnMatch against the previous frame's descriptors, Ransac::Iterate on the 3D-3D pairs —
    host in, host out, a synchronisation after every call.  Milliseconds per frame (frame + its pair), wall clock."""
    ctx = ob.Context(max_frames=2, max_pairs=2, device=local)
    fr = [np.array(f) for f in frames[:n]]; dp = [np.array(d) for d in depths[:n]]      # separate pageable arrays, as a loader hands them over
    prev = None; ts = {"extract": [], "match": [], "ransac": [], "total": []}
    for i in range(n):
        t0 = time.perf_counter()
        ctx.extract_batch(fr[i][None], dp[i][None], slot0=i & 1)
        k, d, xyz = ctx.download_frame(i & 1)
        t1 = time.perf_counter()
        if prev is not None:
            m = ctx.knn_match(prev[0], d, RATIO, CROSS)
            t2 = time.perf_counter()
            ctx.ransac_iterate(prev[1], xyz, m, want_table=False, seed=42 + i)
            t3 = time.perf_counter()
            if i >= 8:                                     # the first calls pay allocation / module load
                ts["extract"].append(t1 - t0); ts["match"].append(t2 - t1); ts["ransac"].append(t3 - t2); ts["total"].append(t3 - t0)
        prev = (d, xyz)
    ctx.close()
    q = lambda a, p: float(np.percentile(np.array(a) * 1e3, p))
    py = {"frames": len(ts["total"]), "p50": q(ts["total"], 50), "p99": q(ts["total"], 99), "mean": float(np.mean(ts["total"]) * 1e3),
          "p50_by_call": {k: q(v, 50) for k, v in ts.items() if k != "total"},
          "calls": "orbf_extract_batch(n=1) + orbf_download_frame, orbf_knn_match, orbf_ransac_iterate through ctypes (numpy allocations included)"}
    out = {"unit": "ms per frame (extract + match with the previous frame + RANSAC), one frame at a time, pageable host in / host out"}
    cpp = latency_cpp(frames, depths, n)
    if cpp is not None:
        out.update(cpp)
        out["calls"] = ("C++ host mirror (include/orbfront_host.hpp, tools/latency_cpp.cpp): Frame::ExtractFeatures, Matcher::KnnMatch(last, cur, cross-check), "
                        "Odometry::Compute (Ransac::Iterate + clouds + composition rule), std::chrono around each call")
        out["python_ctypes"] = py
    else:
        out.update(py)
    return out


def latency_cpp(frames, depths, n):
    """Builds tools/latency_cpp.cpp against the C++ host mirror and runs it on the first n frames; None when no host compiler is there."""
    import shutil
    import struct
    import tempfile
    if not shutil.which("g++"):
        return None
    pkg = ROOT / "adaptive-rgbd-localization-mappig_b200"
    with tempfile.TemporaryDirectory() as td:
        exe = Path(td) / "latency_cpp"
        r = subprocess.run(["g++", "-std=c++17", "-O2", f"-I{ROOT / 'include'}", "-o", str(exe), str(ROOT / "tools" / "latency_cpp.cpp"), f"-L{pkg}",
                            "-lorbfront_b200", f"-Wl,-rpath,{pkg}"], capture_output=True, text=True)
        if r.returncode != 0:
            return None
        raw = Path(td) / "in.raw"
        with open(raw, "wb") as f:
            f.write(struct.pack("3i", n, W, H)); f.write(np.ascontiguousarray(frames[:n]).tobytes()); f.write(np.ascontiguousarray(depths[:n]).tobytes())
        r = subprocess.run([str(exe), str(raw)], capture_output=True, text=True)
        if r.returncode != 0:
            return None
        try:
            return json.loads(r.stdout.strip().splitlines()[-1])
        except Exception:
            return None


def fast_smem_view(F, ms, sm_mhz):
    """FAST against the resource that binds it: the shared-memory data pipe (1 wavefront / clk / SM).  Wavefronts per pixel from the committed
    ncu capture (profiles/r2i_fast_kernel.txt: l1tex__data_pipe_lsu_wavefronts_mem_shared.sum / pixels of the launch); achieved = that x the
    pixels of this launch / its CUDA-event time."""
    wpp = 0.445; px = 950_532 * F; clk = (sm_mhz or 1965.0) * 1e6
    ach = wpp * px / (ms * 1e-3)
    return {"bound": "shared-memory data pipe", "wavefronts_per_pixel": wpp, "achieved_wavefronts_per_s": ach, "peak_wavefronts_per_s": 148 * clk,
            "frac": ach / (148 * clk), "source": "profiles/r2i_fast_kernel.txt (ncu --set full of this kernel at 128 frames)"}


def int_pipe_view(F, stage_ms, sm_mhz):
    """The integer-pipe view of the three stencil stages (north star: '>= 60 % of the HBM or integer-pipe roofline per stage').  Every
    integer pipe of a B200 SM issues 64 lanes / clk (profiles/int_pipe_peaks.json, tools/pipe_bench.cu: LOP3, IADD3, VIMNMX, IDP.4A, IMAD
    all 64 / clk / SM).  algorithmic = thread-level operations the stage's arithmetic needs on its binding pipe per pixel, derived in
    DESIGN.md section 4.  frac = algorithmic ops/s / peak."""
    clk = (sm_mhz or 1965.0) * 1e6
    peak = 64.0 * 148 * clk
    px = 950_532 * F
    rows = {"blur7": ("IMAD / IDP pipe (ncu: math-pipe throttle is the top stall; issue slots 72 %)", 5.25,
                      "7 IDP.4A (28 horizontal MACs / 4) + 14 IDP.2A (28 vertical MACs / 2) per 4 pixels; the kernel issues 10 + 12 + 4 per 4 pixels on 42 rows per 36"),
            "fast_cell": ("ALU pipe (ncu: 57 % busy; the kernel is bound by the shared-memory data pipe, see fast_cell_smem)", 9.6,
                          "pretest 8 VABSDIFF4 + 8 IADD + 8 LOP3 per 4 pixels = 6 / px, strength 80 packed min / max per 2 survivors x 8.3 % = 3.3 / px, NMS 8 compares x 3.6 % = 0.3 / px"),
            "pyr_resize": ("IMAD / IDP pipe (ncu: ALU pipe 48 %, issue slots 68 %, warps active 61 %)", 3.2,
                           "per destination pixel of levels 1..7 (643,332 px / frame): 1.2 source rows x 1 IDP.2A + 2 IMAD.HI")}
    out = {"peak_lane_ops_per_s_per_pipe": peak, "peak_source": "measured 64 lanes / clk / SM for every integer pipe (profiles/int_pipe_peaks.json)"}
    for k, (pipe, alg, how) in rows.items():
        n = (643_332 if k == "pyr_resize" else 950_532) * F
        out[k] = {"pipe": pipe, "algorithmic_ops_per_px": alg, "frac": alg * n / (stage_ms[k] * 1e-3) / peak, "how": how}
    return out


def ransac_view(summ, stage_ms):
    """unit = hypothesis x pair Mahalanobis evaluation (~150 f64 flops).  Counted as the reference's loop would execute them at
    least: one scoring pass over the pair's good matches per VALID hypothesis (the <= 19 refit passes per hypothesis are not counted), over
    the time of the RANSAC kernels of the step.  Latency-bound (a chain of ~10 dependent launches), no roofline claimed."""
    evals = float(np.sum(summ["valid_iters"].astype(np.int64) * summ["n_good"].astype(np.int64)))
    ms = stage_ms["ransac_prepare"] + stage_ms["ransac_hyp"] + stage_ms["ransac_select"]
    return {"hyp_pair_evals_per_step": evals, "ms_per_step": ms, "hyp_pair_evals_per_s": evals / (ms * 1e-3), "f64_flops_per_eval": 150,
            "f64_flops_per_s": 150 * evals / (ms * 1e-3), "bound": "latency (dependent launches; warp per hypothesis, f64 Mahalanobis)",
            "mean_valid_hypotheses_per_pair": float(np.mean(summ["valid_iters"])), "mean_good_matches_per_pair": float(np.mean(summ["n_good"]))}


def knn2_view(frame_counts, ms):
    """The matcher is the one compute-bound stage: descriptor pairs/s, and the same number as int8 tensor-core throughput
    (tcgen05 kind::i8: every pair is a 256-long s8 dot product = 512 ops; dense peak = 16384 ops/clk/SM, the figure ncu
    reports as sm__ops_path_tensor_op_utcimma_src_int8 peak, x 148 SMs x 1.965 GHz) and against the POPC formulation it
    replaced (8 POPC32 per pair at 16 lanes/clk/SM)."""
    pairs = float(sum(int(a) * int(b) for a, b in zip(frame_counts[:-1], frame_counts[1:])))
    pps = pairs / (ms * 1e-3)
    nominal = 16384 * 148 * 1.965e9
    tensor_peak = nominal; src = "nominal (16384 int8 ops / clk / SM)"
    pfile = ROOT / "profiles" / "int8_tensor_peak.json"
    if pfile.exists():                                   # bare tcgen05.mma kind::i8 loop on this pool's B200 (tools/umma_peak.cu)
        try:
            pj = json.loads(pfile.read_text())
            tensor_peak = max(float(pj["int8_ops_per_s_1cta_per_sm"]), float(pj["int8_ops_per_s_2cta_per_sm"]))
            src = "measured (profiles/int8_tensor_peak.json: bare UTCIMMA loop, tools/umma_peak.cu)"
        except Exception:
            pass
    return {"pairs_per_s": pps, "bound": "tensor", "int8_ops_per_pair": 512, "achieved_int8_ops_per_s": pps * 512, "int8_dense_peak_ops_per_s": tensor_peak,
            "peak_source": src, "nominal_int8_dense_peak_ops_per_s": nominal,
            "frac_of_tensor_peak": pps * 512 / tensor_peak, "popc32_per_pair_if_popc": 8, "nominal_popc_ceiling_pairs_per_s": 16 * 148 * 1.965e9 / 8,
            "note": "ncu: tensor-active % + ALU-active % ~ 100 % (the two co-resident CTAs run their MMA and epilogue phases in lockstep), so the kernel "
                    "time is T_tensor + T_alu; the epilogue (top-2 + cross-check butterfly + operand expansion), not the tensor pipe, sets it"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--frames", type=int, default=512, help="frames per step per GPU")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=192, help="frames of the cpu_baseline sample (rank 0, N=1; 0 = skip)")
    ap.add_argument("--chunk", type=int, default=0, help="pipeline chunk in frames (0 = library default, <0 = no chunking)")
    ap.add_argument("--streams", type=int, default=0, help="pipeline worker streams (0 = library default)")
    ap.add_argument("--config", type=int, default=3, choices=[3, 4, 5], help="3 = the headline sequence workload (default); 4 = 1280x720 / 2000 kp "
                    "with adapted thresholds; 5 = keyframe-database many-to-many matching over NCCL / NVLink")
    ap.add_argument("--latency-frames", type=int, default=200, help="frames of the one-frame-at-a-time latency arm (rank 0, N=1; 0 = skip)")
    ap.add_argument("--no-overlap", action="store_true", help="e2e arm: every call starts after the previous one has finished on the device")
    ap.add_argument("--depth-copy", action="store_true", help="e2e arm: stage whole depth planes in HBM instead of sampling pinned host memory in place")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.config == 5:
        return run_config5(args)
    if args.config == 4:
        return run_config4(args)

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        # NCCL prints its version banner on fd 1 while the communicator comes up; the contract is ONE JSON line on stdout, so
        # stdout points at stderr until the first collective has run
        sys.stdout.flush()
        saved = os.dup(1); os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush(); os.dup2(saved, 1); os.close(saved)
    ob = load_pkg()
    F = args.frames
    frames, depths = make_inputs(F, seed=rank)            # each rank its own shard of the sequence (weak scaling)
    try:        # one group of host cores per rank: the ranks' launch threads and pinned-buffer traffic do not migrate onto each other
        ncpu = os.cpu_count() or 1
        per = max(1, ncpu // max(world, 1))
        os.sched_setaffinity(0, set(range(local * per, min(ncpu, (local + 1) * per))) or set(range(ncpu)))
    except (AttributeError, OSError):
        pass
    # device-resident arm: two slot halves of one context, consecutive steps alternate between them so that the RANSAC of step i (a chain of
    # latency-bound launches on the library's side stream) runs under the pyramid / FAST kernels of step i + 1 (pipeline_overlap)
    ctx = ob.Context(max_frames=2 * F, max_pairs=2 * F, device=local, pipeline_chunk=args.chunk, pipeline_streams=args.streams,
                     depth_zero_copy=-1 if args.depth_copy else 0, pipeline_overlap=0 if args.no_overlap else 1)
    stream = torch.cuda.Stream(device=local)
    ctx.set_stream(stream.cuda_stream)
    # the end-to-end arm double-buffers whole sequences in one context (frame / pair slot halves): the H2D copies of step i + 1 run under
    # the last kernels, RANSAC and result read-back of step i
    ectx = ob.Context(max_frames=2 * F, max_pairs=2 * F, device=local, pipeline_chunk=args.chunk, pipeline_streams=args.streams,
                      depth_zero_copy=-1 if args.depth_copy else 0, pipeline_overlap=0 if args.no_overlap else 1)

    d_gray = torch.from_numpy(frames).cuda(local)
    d_depth = torch.from_numpy(depths.view(np.int16)).cuda(local)
    h_gray = torch.from_numpy(frames).pin_memory(); h_depth = torch.from_numpy(depths.view(np.int16)).pin_memory()
    hg = h_gray.numpy(); hd = h_depth.numpy().view(np.uint16)
    torch.cuda.synchronize()

    dev_step = [0]

    def step_device():
        h = dev_step[0] & 1; dev_step[0] += 1
        ctx.track_sequence_device(d_gray.data_ptr(), W, W * H, F, d_depth.data_ptr(), W, W * H, RATIO, CROSS, seed=42, slot0=h * F, pair_slot0=h * F)

    pin = lambda shape, dt: torch.zeros(shape, dtype=dt).pin_memory()
    res = [dict(fc=pin(F, torch.int32), mc=pin(F - 1, torch.int32), rr=pin((F - 1) * ob.RANSAC_RESULT_DT.itemsize, torch.uint8)) for _ in range(2)]
    K = ectx.K
    feat = None

    def issue_e2e(i, full=False):
        """Step i of the end-to-end arm: pinned host frames -> H2D -> path -> asynchronous D2H of the results (poses, inlier / match /
        keypoint counts; with full=True also every keypoint, descriptor, 3D point and match) into pinned host memory."""
        h = i & 1
        ectx.track_sequence_at(hg, hd, RATIO, h * F, h * F, CROSS, seed=42)
        r = res[h]
        if full:
            f = feat[h]
            ectx.read_features_async(h * F, h * F, F, f["kps"].numpy().view(ob.KEYPOINT_DT), f["desc"].numpy(), f["xyz"].numpy(),
                                     f["m"].numpy().view(ob.DMATCH_DT), 2 + h)
        ectx.read_results_async(h * F, h * F, F, r["fc"].numpy(), r["mc"].numpy(), r["rr"].numpy().view(ob.RANSAC_RESULT_DT), h)

    def run_e2e(steps, full=False):
        issue_e2e(0, full)
        for i in range(1, steps):
            issue_e2e(i, full)
            ectx.wait_marker((i - 1) & 1)                   # results of step i - 1 are on the host
        ectx.wait_marker((steps - 1) & 1)
        r = res[(steps - 1) & 1]
        return r["rr"].numpy().view(ob.RANSAC_RESULT_DT).copy(), r["mc"].numpy().copy(), r["fc"].numpy().copy()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing ----
    sampler = ClockSampler(local); sampler.start()             # started early: nvidia-smi needs 0.1-1 s before its first row
    step_device(); torch.cuda.synchronize()
    t_wait = time.perf_counter()
    while sampler.proc and not sampler.rows and time.perf_counter() - t_wait < 5.0:     # keep the GPU busy until the sampler is live
        step_device(); torch.cuda.synchronize()
    t_load0 = time.perf_counter(); sampler.load_t0 = t_load0
    for _ in range(args.warmup):
        step_device()
    barrier()
    l0 = ctx.launch_count()
    ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
    barrier()
    tw0 = time.perf_counter()
    with torch.cuda.stream(stream):
        ev0.record(stream)
        for _ in range(args.steps):
            step_device()
        ctx.join()                                          # the last step's RANSAC (side stream) is part of the timed region
        ev1.record(stream)
    barrier()
    tw1 = time.perf_counter()
    launches = ctx.launch_count() - l0
    ms = ev0.elapsed_time(ev1) / args.steps
    # ---- end-to-end timing (host buffers, copies inside the timed region) ----
    run_e2e(max(2, args.warmup))
    barrier()
    t0 = time.perf_counter()
    summ, mc, fc = run_e2e(args.steps)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    e2e_ms = (t1 - t0) * 1e3 / args.steps
    barrier()
    # the same with everything the reference's calls hand back to host vectors (keypoints, descriptors, 3D points, matches) read back too
    feat = [dict(kps=pin(F * K * ob.KEYPOINT_DT.itemsize, torch.uint8), desc=pin((F, K, 32), torch.uint8), xyz=pin((3, F, K), torch.float32),
                 m=pin((F - 1) * K * ob.DMATCH_DT.itemsize, torch.uint8)) for _ in range(2)]
    run_e2e(2, full=True)
    barrier()
    tf0 = time.perf_counter()
    run_e2e(args.steps, full=True)
    torch.cuda.synchronize()
    tf1 = time.perf_counter()
    full_ms = (tf1 - tf0) * 1e3 / args.steps
    full_d2h = int(sum(t.numel() * t.element_size() for t in feat[0].values()))
    barrier()
    clocks = sampler.stop([(tw0, tw1), (t0, t1)])
    clocks["load_window_s"] = round(t1 - t_load0, 3)
    # depth: either the whole u16 planes are staged, or (pinned host memory, the default) only the 32-byte sectors holding the
    # one sample each keypoint needs cross PCIe, read in place by the unprojection kernel
    h2d = int(frames.nbytes + (depths.nbytes if args.depth_copy else 32 * int(np.sum(fc))))
    d2h = int(summ.nbytes + mc.nbytes + fc.nbytes)
    # ---- per-stage times (CUDA events inside the library, on the launching stream) ----
    ctx.profile_enable(True)
    per_iter = []
    prev = {k: 0.0 for k in ctx.profile_read()}
    for _ in range(max(5, args.steps)):
        step_device(); ctx.profile_collect()
        cur = {k: v[0] for k, v in ctx.profile_read().items()}
        per_iter.append({k: cur[k] - prev[k] for k in cur}); prev = cur
    ctx.profile_enable(False)
    stage_ms = {k: float(np.median([it[k] for it in per_iter])) for k in per_iter[0]}     # median over the profiled steps

    per_rank = None
    if world > 1:
        mine = torch.tensor([ms, e2e_ms, full_ms], device=f"cuda:{local}", dtype=torch.float64)
        allr = torch.empty((world, 3), device=f"cuda:{local}", dtype=torch.float64)
        dist.all_gather_into_tensor(allr, mine)
        per_rank = {"device_ms": [float(x) for x in allr[:, 0]], "e2e_ms": [float(x) for x in allr[:, 1]], "e2e_full_ms": [float(x) for x in allr[:, 2]]}
        ms, e2e_ms, full_ms = float(allr[:, 0].max()), float(allr[:, 1].max()), float(allr[:, 2].max())
    total_frames = F * world
    value = total_frames / (ms * 1e-3)
    e2e_value = total_frames / (e2e_ms * 1e-3)

    if rank == 0:
        peaks_path = ROOT / "MEASURED_PEAKS.json"
        if peaks_path.exists():
            peak = float(json.loads(peaks_path.read_text())["hbm_gbs"]); peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak = 6650.0; peak_src = "fallback (B200_PROFILING.md 6.65 TB/s)"
        hbm_stages = {k: stage_ms[k] for k in STAGE_BYTES}
        dom = max(hbm_stages, key=hbm_stages.get)
        launches_per_step = {"pyr_resize": 7, "fast_cell": 1, "blur7": 1}[dom]
        bytes_per_launch = STAGE_BYTES[dom] * F / launches_per_step
        achieved = STAGE_BYTES[dom] * F / (stage_ms[dom] * 1e-3) / 1e9
        # dram__bytes_read.sum + dram__bytes_write.sum of the dominant stage from the committed ncu --set full capture
        # (profiles/traffic.json: bytes per frame, measured at 128 frames per launch), scaled to this run's launch
        traffic = None; traffic_note = None
        tfile = ROOT / "profiles" / "traffic.json"
        if tfile.exists():
            try:
                tj = json.loads(tfile.read_text())
                traffic = tj[dom]["dram_bytes_per_frame"] * F / launches_per_step
                traffic_note = tj["source"]
            except Exception:
                traffic = None
        roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": traffic, "traffic_source": traffic_note, "algorithmic_bytes_per_launch": bytes_per_launch, "launches_per_step": launches_per_step,
                    "avg_launch_ms": stage_ms[dom] / launches_per_step, "peak_source": peak_src,
                    "note": ("the dominant kernel by time (fast_cell) is bound by the shared-memory data pipe, not by HBM (ncu, profiles/r2i_fast_kernel.txt: "
                             "81 % of the peak wavefront rate, issue slots 71 %, DRAM 6 %): its HBM fraction is small by construction and fast_cell_smem "
                             "is its real roofline.  stage_hbm_frac has the HBM-class stages (pyramid resize, blur), hamming_knn2 the tensor-core view "
                             "of the matcher, ransac the hypothesis x pair rate."),
                    "stage_ms_per_step": stage_ms,
                    "stage_hbm_frac": {k: STAGE_BYTES[k] * F / (stage_ms[k] * 1e-3) / 1e9 / peak for k in STAGE_BYTES},
                    "whole_path": {"algorithmic_bytes_per_frame": FRAME_BYTES, "achieved_GBps": FRAME_BYTES * F / (ms * 1e-3) / 1e9,
                                   "frac_of_hbm": FRAME_BYTES * F / (ms * 1e-3) / 1e9 / peak},
                    "hamming_knn2": knn2_view(fc, stage_ms["hamming_knn2"]),
                    "fast_cell_smem": fast_smem_view(F, stage_ms["fast_cell"], clocks.get("sm_mhz")),
                    "int_pipe": int_pipe_view(F, stage_ms, clocks.get("sm_mhz")),
                    "ransac": ransac_view(summ, stage_ms)}
        cpu = None; parity = None; latency = None
        if world == 1 and args.latency_frames > 0:
            latency = latency_arm(ob, local, frames, depths, min(F, args.latency_frames))
        if world == 1 and args.cpu_sample > 0:
            from oracle import oracle as orc
            orc.build()
            ns = min(args.cpu_sample, F)
            t0 = time.perf_counter()
            cpu_pipeline(orc, frames[:ns], depths[:ns])
            dt = time.perf_counter() - t0
            cpu = {"value": ns / dt, "unit": "frames/s", "cores": 1, "kind": "port",
                   "sample": f"first {ns} frames of the step's batch (extract {ns}, match+RANSAC {ns - 1} pairs), oracle -O3 -march=native, {dt:.1f} s"}
            run_e2e(1); torch.cuda.synchronize()            # the results compared are those of the end-to-end path (slot half 0)
            parity = parity_in_bench(orc, ectx, frames, depths, ns, 42)
            if latency is not None:
                latency["cpu_port_ms_per_frame"] = 1e3 / cpu["value"]
        out = {"metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
               "config": config_object(F),
               "arm": {"depth": "e2e: " + ("whole planes copied to HBM" if args.depth_copy else "pinned host planes sampled in place over PCIe (one 32-byte sector per keypoint); --depth-copy stages them instead"),
                       "pipeline": (f"e2e: orbf_track_sequence_at into alternating slot halves of one context, H2D on a copy stream, chunks of {ectx.cfg.pipeline_chunk or (64 if args.no_overlap else 256)} "
                                    f"frames over {ectx.cfg.pipeline_streams or 4} worker streams, " + ("calls serialised" if args.no_overlap else "consecutive calls overlap (pipeline_overlap): results of "
                                    "step i are read back while step i + 1 runs") + "; device-resident: one stream, "
                                    + ("steps serialised" if args.no_overlap else "steps alternate between two slot halves, the RANSAC of step i runs on the side stream under the pyramid / FAST of step i + 1")),
                       "per_rank_ms": per_rank},
               "clocks": clocks, "e2e": {"value": e2e_value, "unit": "frames/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": h2d,
                                         "d2h_bytes_per_step": d2h},
               "e2e_full_download": {"value": total_frames / (full_ms * 1e-3), "unit": "frames/s", "ms_per_step": full_ms, "h2d_bytes_per_step": h2d,
                                     "d2h_bytes_per_step": d2h + full_d2h,
                                     "what": "as e2e, plus every frame's keypoints (cv::KeyPoint layout), descriptors, 3D points and every pair's matches copied to pinned host memory"},
               "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "parity_in_bench": parity, "latency": latency,
               "results": {"mean_keypoints": float(np.mean(fc)), "mean_matches": float(np.mean(mc)),
                           "ransac_ok_frac": float(np.mean(summ["ok"])), "mean_inliers": float(np.mean(summ["n_inliers"]))}}
        print(json.dumps(out))
        if parity is not None and not parity["identical"]:
            ctx.close(); ectx.close()
            raise SystemExit("bench.py: the timed path disagrees with the oracle: " + "; ".join(parity["mismatches"]))
    ctx.close(); ectx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
