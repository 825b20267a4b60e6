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
MODE = os.environ.get("H2D_HOST", "pinned")          # pinned (cudaHostAlloc default, what torch's pin_memory gives) | wc (write-combined)
d = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
if MODE == "wc":
    import ctypes as C
    rt = C.CDLL("libcudart.so.12")
    hp = C.c_void_p()
    assert rt.cudaHostAlloc(C.byref(hp), C.c_size_t(nbytes), C.c_uint(4)) == 0          # cudaHostAllocWriteCombined
    C.memset(hp, 7, nbytes)
    stream = torch.cuda.current_stream().cuda_stream
    class H:
        pass
    def copy():
        assert rt.cudaMemcpyAsync(C.c_void_p(d.data_ptr()), hp, C.c_size_t(nbytes), C.c_int(1), C.c_void_p(stream)) == 0
else:
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory(); h.fill_(7)
    def copy():
        d.copy_(h, non_blocking=True)
for _ in range(3): copy()
torch.cuda.synchronize()
if world > 1: dist.barrier()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(20): copy()
torch.cuda.synchronize(); ms = (time.perf_counter() - t0) / 20 * 1e3
if world > 1:
    t = torch.tensor([ms], device="cuda", dtype=torch.float64); allt = torch.empty(world, device="cuda", dtype=torch.float64)
    dist.all_gather_into_tensor(allt, t); per_rank = [float(x) for x in allt]
else:
    per_rank = [ms]
if rank == 0:
    print(json.dumps({"n_gpus": world, "host_memory": MODE, "bytes_per_copy": nbytes, "ms_per_copy_per_rank": per_rank, "gbs_per_rank": [nbytes / m / 1e6 for m in per_rank],
                      "aggregate_gbs": sum(nbytes / m / 1e6 for m in per_rank), "host_cpus": os.cpu_count()}))
if world > 1: dist.destroy_process_group()
