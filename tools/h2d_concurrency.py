#!/usr/bin/env python3
"""What N concurrent plain H2D copies achieve on this box: every rank copies the bench step's gray planes (157 MB, pinned) to its GPU
20 times after a barrier; rank 0 prints per-rank ms / GB/s.  Launch: python -m torch.distributed.run --nproc-per-node N tools/h2d_concurrency.py
(N = 1 works without torchrun).  The end-to-end arm of bench.py cannot go faster than this."""
import json, os, time
import torch, torch.distributed as dist
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local)); dist.barrier()
try:
    ncpu = os.cpu_count() or 1; per = max(1, ncpu // world)
    os.sched_setaffinity(0, set(range(local * per, min(ncpu, (local + 1) * per))))
except (AttributeError, OSError):
    pass
nbytes = 512 * 640 * 480
h = torch.empty(nbytes, dtype=torch.uint8).pin_memory(); h.fill_(7)
d = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
for _ in range(3): d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
if world > 1: dist.barrier()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(20): d.copy_(h, non_blocking=True)
torch.cuda.synchronize(); ms = (time.perf_counter() - t0) / 20 * 1e3
if world > 1:
    t = torch.tensor([ms], device="cuda", dtype=torch.float64); allt = torch.empty(world, device="cuda", dtype=torch.float64)
    dist.all_gather_into_tensor(allt, t); per_rank = [float(x) for x in allt]
else:
    per_rank = [ms]
if rank == 0:
    print(json.dumps({"n_gpus": world, "bytes_per_copy": nbytes, "ms_per_copy_per_rank": per_rank, "gbs_per_rank": [nbytes / m / 1e6 for m in per_rank],
                      "aggregate_gbs": sum(nbytes / m / 1e6 for m in per_rank), "host_cpus": os.cpu_count()}))
if world > 1: dist.destroy_process_group()
