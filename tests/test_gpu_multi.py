"""Multi-process GPU tests (skipped with fewer than 2 GPUs): the config-5 exchange step through the C ABI — ncclAllGather of the keyframe
stores (orbf_comm_* / orbf_kfdb_allgather, NCCL loaded with dlopen) and the collective-free route (CUDA IPC peer stores read over
NVLink by the matcher, orbf_kfdb_attach_peers) — both against the oracle on every rank."""
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    sys.path[:0] = [str(ROOT), str(ROOT / "tests")]
    import synth
    from conftest import load_orbfront
    from oracle import oracle as orc
    ob = load_orbfront(); orc.build()
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        tex = synth.make_texture(0, 480, 640)
        kf_local = 5
        frames = np.stack([synth.make_frame(tex, 10 * rank + i) for i in range(kf_local)])
        q = synth.make_frame(tex, 77 + rank)[None]
        ctx = ob.Context(max_frames=kf_local + 1, device=rank)
        ctx.extract_batch(frames); ctx.extract_batch(q, slot0=kf_local)
        nq = int(ctx.frame_counts(1, slot0=kf_local)[0])
        ctx.kfdb_reserve(kf_local)
        for k in range(kf_local):
            ctx.kfdb_add_from_slot(k, k)
        idt = torch.from_numpy(ob.comm_unique_id() if rank == 0 else np.zeros(128, np.uint8)).cuda(rank)
        dist.broadcast(idt, 0)
        ctx.comm_init(idt.cpu().numpy(), world, rank)
        total = ctx.kfdb_allgather()
        assert total == world * kf_local
        ctx.kfdb_match_slot(kf_local, 0, total, 0.8)
        got_gather = ctx.kfdb_results(total, nq)
        a, b = ctx.kfdb_ipc_handles()
        allh = torch.empty((world, 128), dtype=torch.uint8, device=f"cuda:{rank}")
        dist.all_gather_into_tensor(allh, torch.from_numpy(np.concatenate([a, b])).cuda(rank))
        hh = allh.cpu().numpy()
        ctx.kfdb_attach_peers(hh[:, :64].copy(), hh[:, 64:].copy(), world, rank, kf_local)
        ctx.kfdb_match_slot(kf_local, 0, total, 0.8)
        got_peers = ctx.kfdb_results(total, nq)
        qd = orc.extract(q[0])[1]
        for g in range(total):
            r, j = divmod(g, kf_local)
            td = orc.extract(synth.make_frame(tex, 10 * r + j))[1]
            ref = orc.knn2(qd, td); ns = len(orc.knn_match(qd, td, 0.8))
            for name, got in (("allgather", got_gather), ("peers", got_peers)):
                assert all(np.array_equal(got[k][g], ref[k]) for k in range(4)) and int(got[4][g]) == ns, (name, rank, g)
        dist.barrier()
        ctx.kfdb_detach_peers(); ctx.comm_destroy(); ctx.close()
        (Path(out_dir) / f"ok{rank}").write_text("ok")
    finally:
        dist.destroy_process_group()


def test_config5_allgather_and_peer_routes_two_ranks(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()
