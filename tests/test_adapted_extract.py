"""BASELINE config 4, 8-level variant: the ORB extractor with iniThFAST adapted per image region by stateful controllers (a north-star
extension defined by the oracle; reference pieces: detectoradjuster.cpp:22-59 for the controller arithmetic,
videodynamicadaptedfeaturedetector.cpp:24-44 for "state carries from frame to frame", orbextractor.cpp:665-723 for the cells).
CPU: what the definition implies, on the oracle.  GPU: orbf_extract_adapted against the oracle — keypoints, descriptors, the thresholds
every frame was detected with, the keypoints found per region and the controller state, over clips that drive thresholds both ways."""
import numpy as np
import pytest

import synth


def _clip(n, w=640, h=480, seed=0):
    """Texture clip whose left third loses contrast from frame 3 on (controllers there step down to the minThFAST floor) while the
    rest stays busy (controllers step up)."""
    tex = synth.make_texture(seed, h, w)
    out = []
    for i in range(n):
        img = synth.make_frame(tex, i, w, h, seed)
        if i >= 3:
            img[:, : w // 3] = (img[:, : w // 3].astype(np.float32) * 0.15 + 100).astype(np.uint8)
        out.append(img)
    return np.stack(out)


def test_default_state_on_first_frame_is_the_plain_extractor(orc):
    img = _clip(1)[0]
    th = np.zeros(9)
    k, d, used, found = orc.extract_adapted(img, th)
    k0, d0 = orc.extract(img)
    assert k.tobytes() == k0.tobytes() and np.array_equal(d, d0) and np.all(used == 20) and found.sum() == len(k)
    # one controller step per region: x 0.7 below the band, x 1.3 above it, unchanged inside
    cfg = orc.adaptive_default()
    want = np.where(found < cfg.min_features, 20 * 0.7, np.where(found > cfg.max_features, 20 * 1.3, 20.0))
    assert np.array_equal(th, want)


def test_thresholds_move_both_ways_and_clamp(orc):
    frames = _clip(14)
    th = np.zeros(9)
    lows, highs = [], []
    for f in frames:
        _, _, used, found = orc.extract_adapted(f, th)
        lows.append(used[[0, 3, 6]].min()); highs.append(used.max())
        assert np.all(used >= 7) and np.all(used <= 254)            # never below minThFAST: the empty-cell fallback stays meaningful
    assert lows[-1] == 7 and highs[-1] > 60, (lows, highs)
    assert th.min() >= 2.0 and th.max() <= 10000.0


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,nf,n", [(640, 480, 1000, 12), (1280, 720, 2000, 5)])
def test_cuda_matches_oracle_over_a_clip(ob, orc, w, h, nf, n):
    frames = _clip(n, w, h, seed=3)
    cfg = orc.adaptive_default()
    band = dict(min_features=round(0.6 * nf / 9), max_features=round(1.02 * nf / 9))
    cfg.min_features, cfg.max_features = band["min_features"], band["max_features"]
    ctx = ob.Context(width=w, height=h, nfeatures=nf, max_frames=n)
    try:
        th_ref = np.zeros(9); th_got = np.zeros(9)
        ref = [orc.extract_adapted(f, th_ref, nfeatures=nf, cfg=cfg) for f in frames]
        # the clip in two calls: the state handed back by the first continues the video in the second
        cut = n // 2
        used_a, found_a = ctx.extract_adapted(frames[:cut], th_got, slot0=0, **band)
        used_b, found_b = ctx.extract_adapted(frames[cut:], th_got, slot0=cut, **band)
        used = np.concatenate([used_a, used_b]); found = np.concatenate([found_a, found_b])
        for i in range(n):
            k, d, _ = ctx.download_frame(i)
            assert np.array_equal(used[i], ref[i][2]), f"frame {i}: thresholds used"
            assert k.tobytes() == ref[i][0].tobytes() and np.array_equal(d, ref[i][1]), f"frame {i}: keypoints / descriptors"
            assert np.array_equal(found[i], ref[i][3]), f"frame {i}: keypoints found per region"
        assert np.array_equal(th_got, th_ref), "controller state after the clip"
        assert len({tuple(u) for u in used}) > 3, "the clip must move the thresholds"
    finally:
        ctx.close()


@pytest.mark.gpu
def test_adapted_extraction_leaves_the_plain_path_untouched(ob, orc):
    """The same context afterwards extracts with the global iniThFAST again."""
    frames = _clip(4, seed=5)
    ctx = ob.Context(max_frames=4)
    try:
        th = np.full(9, 90.0)
        ctx.extract_adapted(frames, th)
        ctx.extract_batch(frames)
        for i in range(4):
            k, d, _ = ctx.download_frame(i)
            k0, d0 = orc.extract(frames[i])
            assert k.tobytes() == k0.tobytes() and np.array_equal(d, d0)
    finally:
        ctx.close()


@pytest.mark.gpu
def test_several_videos_advance_together(ob, orc):
    """orbf_extract_adapted_videos: V independent videos, V frames per link of the FAST -> quadtree -> controller chain; every video must
    come out exactly as if it had been run alone (frame t of video v in slot t * V + v)."""
    V, T = 3, 5
    videos = np.stack([_clip(T, seed=10 + v) for v in range(V)])
    videos[1, :, :, 320:] = 90                                   # video 1: right half flat -> its right-hand controllers run down to the floor
    ctx = ob.Context(max_frames=V * T)
    try:
        th = np.zeros((V, 9))
        used, found = ctx.extract_adapted_videos(videos, th)
        for v in range(V):
            th_ref = np.zeros(9)
            for t in range(T):
                k0, d0, u0, f0 = orc.extract_adapted(videos[v, t], th_ref)
                k, d, _ = ctx.download_frame(t * V + v)
                assert np.array_equal(used[t, v], u0) and np.array_equal(found[t, v], f0), f"video {v} frame {t}: thresholds / counts"
                assert k.tobytes() == k0.tobytes() and np.array_equal(d, d0), f"video {v} frame {t}: keypoints / descriptors"
            assert np.array_equal(th[v], th_ref), f"video {v}: controller state"
        assert len({tuple(used[T - 1, v]) for v in range(V)}) == V, "the videos must end on different thresholds"
        # and the single-video entry point still gives the same for video 0
        th1 = np.zeros(9)
        ctx.extract_adapted(videos[0], th1)
        assert np.array_equal(th1, th[0])
        assert ctx.download_frame(T - 1)[0].tobytes() == orc.extract_adapted(videos[0, T - 1], np.array(_state_before_last(orc, videos[0])))[0].tobytes()
    finally:
        ctx.close()


def _state_before_last(orc, clip):
    th = np.zeros(9)
    for f in clip[:-1]:
        orc.extract_adapted(f, th)
    return th
