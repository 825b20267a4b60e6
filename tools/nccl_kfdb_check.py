#!/usr/bin/env python3
"""BASELINE config 5 on N GPUs (torchrun, one rank per GPU): every rank extracts its own keyframes into its shard of the
keyframe store, the shards are all-gathered over NCCL (NVLink) into one device buffer, the matcher attaches to it and
matches a 1000-keypoint query against every keyframe; rank 0 checks the result against the oracle and prints timings.
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/nccl_kfdb_check.py [--kf 64]
"""
import argparse
import importlib.util
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "tests")]
import synth  # noqa: E402


def load(name, rel):
    spec = importlib.util.spec_from_file_location(name, ROOT / "adaptive-rgbd-localization-mappig_b200" / rel)
    mod = importlib.util.module_from_spec(spec); sys.modules[name] = mod; spec.loader.exec_module(mod)
    return mod


def main():
    ap = argparse.ArgumentParser(); ap.add_argument("--kf", type=int, default=32, help="keyframes per rank"); args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ob = load("orbfront_b200", "__init__.py"); sh = load("orbf_sharding", "sharding.py")
    tex = synth.make_texture(0, 480, 640)
    kf = args.kf
    frames = np.stack([synth.make_frame(tex, 2 * (rank * kf + i)) for i in range(kf)])       # every 2nd frame is a keyframe
    ctx = ob.Context(max_frames=kf, device=local)
    ctx.extract_batch(frames)
    ctx.kfdb_reserve(kf)
    for i in range(kf):
        ctx.kfdb_add_from_slot(i, i)
    ctx.synchronize()
    d_ptr, c_ptr, rows, nkf = ctx.kfdb_device_buffers()
    local_desc = sh.device_tensor(d_ptr, (nkf, rows, 32)); local_counts = sh.device_tensor(c_ptr, (nkf,), "i4")
    torch.cuda.synchronize(); dist.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    gd, gc = sh.gather_keyframes(local_desc, local_counts)
    ev1.record(); torch.cuda.synchronize()
    gather_ms = ev0.elapsed_time(ev1)
    ctx.kfdb_attach_device(gd.data_ptr(), gc.data_ptr(), world * kf)
    q_img = synth.make_frame(tex, 2 * (world * kf) + 1)
    qctx = ob.Context(max_frames=1, device=local)
    qk, q = qctx.extract(q_img)
    t0 = time.perf_counter()
    i1, d1, i2, d2, surv = ctx.kfdb_match(q, 0, world * kf, 0.8)
    match_ms = (time.perf_counter() - t0) * 1e3
    ok = True
    if rank == 0:
        from oracle import oracle as orc
        orc.build()
        for k in sorted(set([0, 1, kf - 1, kf, world * kf - 1])):
            r, i = divmod(k, kf)
            dref = orc.extract(synth.make_frame(tex, 2 * (r * kf + i)))[1]
            ref = orc.knn2(q, dref)
            ok &= bool(np.array_equal(i1[k], ref[0]) and np.array_equal(d1[k], ref[1]) and np.array_equal(i2[k], ref[2]) and np.array_equal(d2[k], ref[3]))
            ok &= int(surv[k]) == len(orc.knn_match(q, dref, 0.8))
        pairs = float(len(q)) * float(gc.sum().item())
        print({"world": world, "keyframes": world * kf, "gathered_MB": gd.numel() / 1e6, "all_gather_ms": gather_ms,
               "all_gather_GBps_per_rank": gd.numel() * (world - 1) / world / (gather_ms * 1e-3) / 1e9,
               "match_ms_incl_d2h": match_ms, "descriptor_pairs": pairs, "parity_vs_oracle": ok})
    dist.barrier()
    ctx.close(); qctx.close()
    dist.destroy_process_group()
    if rank == 0 and not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
