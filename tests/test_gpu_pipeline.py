"""GPU parity of the device-resident batch path: extract -> match consecutive pairs -> RANSAC, everything
staying in HBM between stages, vs the oracle chained on the host."""
import numpy as np
import pytest

import synth

pytestmark = pytest.mark.gpu


def test_sequence_extract_match_ransac(ob, orc, texture):
    n = 6
    frames = np.stack([synth.make_frame(texture, i) for i in range(n)])
    depths = np.stack([synth.make_depth(i) for i in range(n)])
    ctx = ob.Context(max_frames=n)
    try:
        ctx.extract_batch(frames, depths)
        pairs = np.array([[i, i + 1] for i in range(n - 1)], np.int32)
        ctx.match_pairs(pairs, 0.8, cross_check=True)
        ctx.ransac_pairs(n - 1, seed=42)
        host = []
        for i in range(n):
            k, d = orc.extract(frames[i])
            xyz, _ = orc.unproject(k, depths[i])
            host.append((k, d, xyz))
        cov = -1.0
        summary = ctx.download_ransac_summary(n - 1)
        for p in range(n - 1):
            (k1, d1, x1), (k2, d2, x2) = host[p], host[p + 1]
            m_ref = orc.knn_match(d1, d2, 0.8, True)
            m_got = ctx.download_matches(p)
            assert m_got.tobytes() == m_ref.tobytes(), f"pair {p}: matches differ"
            i1, e1, i2, e2 = ctx.download_knn(p)
            r1 = orc.knn2(d1, d2)
            assert np.array_equal(i1, r1[0]) and np.array_equal(e1, r1[1]) and np.array_equal(i2, r1[2]) and np.array_equal(e2, r1[3])
            # quirk Q7: the depth covariance is latched by the first scored pair (pair 0) and reused afterwards
            r = orc.ransac_iterate(x1, x2, m_ref, seed=42 + p, depth_cov=cov)
            cov = r["depth_cov"]
            g = ctx.download_ransac(p)
            assert g["ok"] == r["ok"] and g["n_good"] == r["n_good"]
            assert g["inliers"].tobytes() == r["inliers"].tobytes(), f"pair {p}: inlier sets differ"
            assert np.abs(g["T12"] - r["T12"]).max() <= 1e-5 and g["rmse"] == r["rmse"]
            assert g["depth_cov"] == r["depth_cov"]
            assert summary[p]["n_inliers"] == len(r["inliers"])
            assert r["ok"], "synthetic rigid motion must be recovered"
        assert ctx.launch_count() > 0
    finally:
        ctx.close()


@pytest.mark.parametrize("chunk,streams", [(2, 3), (3, 2), (4, 1), (-1, 3)])
def test_track_sequence_pipelined_equals_staged(ob, orc, texture, chunk, streams):
    """orbf_track_sequence (chunked over worker streams, pairs straddling chunks, depth covariance latched by pair 0)
    gives the same bytes as extract_batch + match_pairs + ransac_pairs on one stream, and as the oracle."""
    n = 9
    frames = np.stack([synth.make_frame(texture, 40 + i) for i in range(n)])
    depths = np.stack([synth.make_depth(40 + i) for i in range(n)])
    ref = ob.Context(max_frames=n, pipeline_chunk=-1)
    ctx = ob.Context(max_frames=n, pipeline_chunk=chunk, pipeline_streams=streams)
    try:
        ref.extract_batch(frames, depths)
        pairs = np.array([[i, i + 1] for i in range(n - 1)], np.int32)
        ref.match_pairs(pairs, 0.8, cross_check=True)
        ref.ransac_pairs(n - 1, seed=7)
        for rep in range(2):                      # second pass: worker streams / events are reused, covariance stays latched
            assert ctx.track_sequence(frames, depths, 0.8, cross_check=True, seed=7) == n - 1
            for s in range(n):
                a, b = ctx.download_frame(s), ref.download_frame(s)
                assert all(x.tobytes() == y.tobytes() for x, y in zip(a, b)), f"frame {s}"
            for p in range(n - 1):
                assert ctx.download_matches(p).tobytes() == ref.download_matches(p).tobytes(), f"pair {p}"
                g, r = ctx.download_ransac(p), ref.download_ransac(p)
                assert g["inliers"].tobytes() == r["inliers"].tobytes() and g["T12"].tobytes() == r["T12"].tobytes()
                assert g["rmse"] == r["rmse"] and g["depth_cov"] == r["depth_cov"] and g["real_iters"] == r["real_iters"]
        k0, d0 = orc.extract(frames[0]); k1, d1 = orc.extract(frames[1])
        m = orc.knn_match(d0, d1, 0.8, True)
        assert ctx.download_matches(0).tobytes() == m.tobytes()
        r = orc.ransac_iterate(orc.unproject(k0, depths[0])[0], orc.unproject(k1, depths[1])[0], m, seed=7)
        assert ctx.download_ransac(0)["inliers"].tobytes() == r["inliers"].tobytes()
    finally:
        ctx.close(); ref.close()


def test_device_input_path(ob, orc, texture):
    """Frames already resident in HBM (torch tensors): extract_batch_device reads level 0 in place."""
    import torch
    n = 3
    frames = np.stack([synth.make_frame(texture, 20 + i) for i in range(n)])
    depths = np.stack([synth.make_depth(20 + i) for i in range(n)])
    dg = torch.from_numpy(frames).cuda(); dd = torch.from_numpy(depths.view(np.int16)).cuda()
    ctx = ob.Context(max_frames=n)
    try:
        torch.cuda.synchronize()
        ctx.extract_batch_device(dg.data_ptr(), 640, 640 * 480, n, dd.data_ptr(), 640, 640 * 480)
        for s in range(n):
            k, d, xyz = ctx.download_frame(s)
            ko, do = orc.extract(frames[s])
            assert k.tobytes() == ko.tobytes() and np.array_equal(d, do)
            assert np.array_equal(xyz, orc.unproject(ko, depths[s])[0])
    finally:
        ctx.close()


def test_keyframe_store_many_to_many(ob, orc, texture):
    n = 4
    frames = np.stack([synth.make_frame(texture, 30 + 2 * i) for i in range(n)])
    ctx = ob.Context(max_frames=n)
    try:
        ctx.extract_batch(frames)
        ctx.kfdb_reserve(n + 1)
        descs = []
        for s in range(n):
            ctx.kfdb_add_from_slot(s, s)
            descs.append(orc.extract(frames[s])[1])
        extra = synth.descriptor_sets(500, seed=77)[0]
        ctx.kfdb_add_host(n, extra); descs.append(extra)
        q = orc.extract(synth.make_frame(texture, 33))[1]
        i1, d1, i2, d2, surv = ctx.kfdb_match(q, 0, n + 1, 0.8)
        for k in range(n + 1):
            r = orc.knn2(q, descs[k])
            assert np.array_equal(i1[k], r[0]) and np.array_equal(d1[k], r[1]) and np.array_equal(i2[k], r[2]) and np.array_equal(d2[k], r[3])
            assert surv[k] == len(orc.knn_match(q, descs[k], 0.8))
    finally:
        ctx.close()


def _load_sharding():
    import importlib.util
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    spec = importlib.util.spec_from_file_location("orbf_sharding", root / "adaptive-rgbd-localization-mappig_b200" / "sharding.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_sharded_sequence_equals_whole_sequence(ob, texture):
    """Two ranks' shards (run one after the other on this GPU: halo frame, global pair seeds, broadcast depth covariance)
    reproduce the single-context run of the whole sequence byte for byte."""
    sh = _load_sharding()
    n = 7
    frames = np.stack([synth.make_frame(texture, 60 + i) for i in range(n)])
    depths = np.stack([synth.make_depth(60 + i) for i in range(n)])
    whole = ob.Context(max_frames=n)
    try:
        _, ref, cov = sh.run_sequence_shard(whole, frames, depths, n, 0, 1, seed=11)
    finally:
        whole.close()
    got = []
    for rank in range(2):
        s = sh.frame_shard(n, 2, rank)
        ctx = ob.Context(max_frames=n)
        try:
            _, res, c = sh.run_sequence_shard(ctx, frames[s["first"]:s["stop"]], depths[s["first"]:s["stop"]], n, rank, 2, seed=11,
                                              first_pair_cov=None if rank == 0 else cov)
            assert c == cov
            got += res
        finally:
            ctx.close()
    assert [r["pair"] for r in got] == list(range(n - 1))
    for g, r in zip(got, ref):
        assert g["inliers"].tobytes() == r["inliers"].tobytes() and g["T12"].tobytes() == r["T12"].tobytes() and g["rmse"] == r["rmse"]


def test_keyframe_shards_gathered_into_one_store(ob, orc, texture):
    """Config 5: the per-rank keyframe shard is viewed as a torch tensor (zero copy), 'gathered' (here: two shards
    concatenated on one GPU, the layout NCCL all_gather_into_tensor produces) and attached as an external store."""
    import torch
    sh = _load_sharding()
    frames = np.stack([synth.make_frame(texture, 70 + 3 * i) for i in range(4)])
    ctx = ob.Context(max_frames=4)
    try:
        ctx.extract_batch(frames)
        ctx.kfdb_reserve(2)
        shards, descs = [], []
        for r in range(2):
            for j in range(2):
                ctx.kfdb_add_from_slot(j, 2 * r + j)
                descs.append(orc.extract(frames[2 * r + j])[1])
            ctx.synchronize()
            d_ptr, c_ptr, rows, nkf = ctx.kfdb_device_buffers()
            shards.append((sh.device_tensor(d_ptr, (nkf, rows, 32)).clone(), sh.device_tensor(c_ptr, (nkf,), "i4").clone()))
        gd = torch.cat([s[0] for s in shards]).contiguous(); gc = torch.cat([s[1] for s in shards]).contiguous()
        torch.cuda.synchronize()
        ctx.kfdb_attach_device(gd.data_ptr(), gc.data_ptr(), 4)
        q = orc.extract(synth.make_frame(texture, 75))[1]
        i1, d1, i2, d2, surv = ctx.kfdb_match(q, 0, 4, 0.8)
        for k in range(4):
            r = orc.knn2(q, descs[k])
            assert np.array_equal(i1[k], r[0]) and np.array_equal(d1[k], r[1]) and np.array_equal(i2[k], r[2]) and np.array_equal(d2[k], r[3])
            assert surv[k] == len(orc.knn_match(q, descs[k], 0.8))
        ctx.kfdb_attach_device(0, 0, 0)
    finally:
        ctx.close()


def test_depth_sampled_in_place_from_pinned_memory_equals_staged_copy(ob, texture):
    """orbf_config.depth_zero_copy: with page-locked host depth planes the unprojection reads its samples over PCIe instead
    of staging the planes in HBM; results must be the same bytes as the staged path (and as pageable input)."""
    import torch
    n = 6
    frames = np.stack([synth.make_frame(texture, 80 + i) for i in range(n)])
    depths = np.stack([synth.make_depth(80 + i) for i in range(n)])
    pg = torch.from_numpy(frames).pin_memory(); pd = torch.from_numpy(depths.view(np.int16)).pin_memory()
    hg = pg.numpy(); hd = pd.numpy().view(np.uint16)
    staged = ob.Context(max_frames=n, depth_zero_copy=-1, pipeline_chunk=2)
    inplace = ob.Context(max_frames=n, pipeline_chunk=2)
    try:
        staged.track_sequence(hg, hd, 0.8, cross_check=True, seed=3)
        inplace.track_sequence(hg, hd, 0.8, cross_check=True, seed=3)
        for s in range(n):
            a, b = inplace.download_frame(s), staged.download_frame(s)
            assert all(x.tobytes() == y.tobytes() for x, y in zip(a, b)), f"frame {s}"
            assert np.any(a[2][:, 2] > 0), "depth must have been sampled"
        for p in range(n - 1):
            g, r = inplace.download_ransac(p), staged.download_ransac(p)
            assert g["inliers"].tobytes() == r["inliers"].tobytes() and g["T12"].tobytes() == r["T12"].tobytes()
        inplace.extract_batch(frames, depths)                       # pageable input on the same context: falls back to staging
        assert all(x.tobytes() == y.tobytes() for x, y in zip(inplace.download_frame(1), staged.download_frame(1)))
    finally:
        staged.close(); inplace.close()


def test_device_steps_overlap_on_alternating_slot_halves(ob, texture):
    """orbf_track_sequence_device_at under pipeline_overlap: the RANSAC of a call runs on the side stream and is joined lazily, the next
    call (other slot half) starts under it.  Several back-to-back calls without any synchronisation must leave exactly the results
    of the serial context in both halves; a call on the SAME half must order itself behind the pending RANSAC."""
    import torch
    n = 6
    seqs = []
    for base in (40, 70):
        fr = np.stack([synth.make_frame(texture, base + i) for i in range(n)]); dp = np.stack([synth.make_depth(base + i) for i in range(n)])
        seqs.append((fr, dp, torch.from_numpy(fr).cuda(), torch.from_numpy(dp.view(np.int16)).cuda()))
    torch.cuda.synchronize()
    ref = []
    plain = ob.Context(max_frames=n)
    try:
        for fr, dp, dg, dd in seqs:
            plain.track_sequence_device(dg.data_ptr(), 640, 640 * 480, n, dd.data_ptr(), 640, 640 * 480, 0.8, True, seed=42)
            plain.synchronize()
            ref.append(([plain.download_frame(s) for s in range(n)], [plain.download_matches(p) for p in range(n - 1)], [plain.download_ransac(p) for p in range(n - 1)]))
    finally:
        plain.close()
    ctx = ob.Context(max_frames=2 * n, max_pairs=2 * n, pipeline_overlap=1)
    try:
        for rep in range(4):                                     # A0 B1 A0 B1 ... : each call overlaps the previous call's RANSAC
            for h, (fr, dp, dg, dd) in enumerate(seqs):
                ctx.track_sequence_device(dg.data_ptr(), 640, 640 * 480, n, dd.data_ptr(), 640, 640 * 480, 0.8, True, seed=42, slot0=h * n, pair_slot0=h * n)
        ctx.join()
        ctx.synchronize()
        for h in range(2):
            frames_ref, matches_ref, ransac_ref = ref[h]
            for s in range(n):
                k, d, xyz = ctx.download_frame(h * n + s)
                assert k.tobytes() == frames_ref[s][0].tobytes() and np.array_equal(d, frames_ref[s][1]) and np.array_equal(xyz, frames_ref[s][2])
            for p in range(n - 1):
                assert ctx.download_matches(h * n + p).tobytes() == matches_ref[p].tobytes(), f"half {h} pair {p}: matches"
                g, r = ctx.download_ransac(h * n + p), ransac_ref[p]
                assert g["ok"] == r["ok"] and g["inliers"].tobytes() == r["inliers"].tobytes() and np.array_equal(g["T12"], r["T12"]), f"half {h} pair {p}: RANSAC"
        # the same half twice in a row: the second call must wait for the first one's RANSAC before it overwrites the pair slots
        fr, dp, dg, dd = seqs[1]
        ctx.track_sequence_device(dg.data_ptr(), 640, 640 * 480, n, dd.data_ptr(), 640, 640 * 480, 0.8, True, seed=42, slot0=0, pair_slot0=0)
        fr, dp, dg, dd = seqs[0]
        ctx.track_sequence_device(dg.data_ptr(), 640, 640 * 480, n, dd.data_ptr(), 640, 640 * 480, 0.8, True, seed=42, slot0=0, pair_slot0=0)
        for p in range(n - 1):
            g, r = ctx.download_ransac(p), ref[0][2][p]
            assert g["inliers"].tobytes() == r["inliers"].tobytes() and np.array_equal(g["T12"], r["T12"])
    finally:
        ctx.close()
