"""Adaptive-threshold FAST detector route (BASELINE config 4): Extractor(FAST, ., ADAPTIVE) =
VideoGridAdaptedFeatureDetector(3x3) over VideoDynamicAdaptedFeatureDetector over DetectorAdjuster(FAST).

  not gpu : the oracle's restatement against the same flow driven through cv2.FastFeatureDetector on sub-image views
            (the call the reference makes, detectoradjuster.cpp:26-27), over a clip so the threshold state evolves;
  gpu     : orbf_adaptive_detect (one response plane + histogram lookups) against the oracle: keypoints, per-cell
            thresholds and counts identical, state carried across calls and across internal sub-batches."""
import numpy as np
import pytest

import synth


def cv2_adaptive(cv2, img, thresh, cfg, retain_best):
    """The reference flow through cv2 (Features/video*adaptedfeaturedetector.cpp), tie rules as documented in the oracle."""
    h, w = img.shape
    g = cfg.grid
    allk = []
    used = np.zeros(g * g, np.int32); found = np.zeros(g * g, np.int32)
    for i in range(g):
        r0, r1 = max((i * h) // g - cfg.edge, 0), min(h, ((i + 1) * h) // g + cfg.edge)
        for j in range(g):
            c0, c1 = max((j * w) // g - cfg.edge, 0), min(w, ((j + 1) * w) // g + cfg.edge)
            sub = img[r0:r1, c0:c1]
            it = cfg.max_iters
            while True:
                t = int(thresh[i * g + j])
                kps = cv2.FastFeatureDetector_create(t, True).detect(sub)
                n = len(kps)
                if n < cfg.min_features:
                    thresh[i * g + j] = max(thresh[i * g + j] * cfg.dec, cfg.min_th)
                elif n > cfg.max_features:
                    thresh[i * g + j] = min(thresh[i * g + j] * cfg.inc, cfg.max_th)
                    break
                else:
                    break
                it -= 1
                if not (it > 0 and cfg.min_th < thresh[i * g + j] < cfg.max_th):
                    break
            used[i * g + j] = t; found[i * g + j] = n
            if n > cfg.max_per_cell:
                order = sorted(range(n), key=lambda a: (-kps[a].response, a))[:cfg.max_per_cell]
                kps = [kps[a] for a in sorted(order)]
            allk += [(k.pt[0] + c0, k.pt[1] + r0, k.size, k.angle, k.response, k.octave, k.class_id) for k in kps]
    if retain_best and len(allk) > retain_best:
        cut = sorted((k[4] for k in allk), reverse=True)[retain_best - 1]
        allk = [k for k in allk if k[4] >= cut]
    return allk, used, found


@pytest.mark.parametrize("w,h", [(640, 480), (1280, 720)])
def test_oracle_adaptive_route_equals_cv2_flow(orc, w, h):
    cv2 = pytest.importorskip("cv2")
    tex = synth.make_texture(3, h, w)
    cfg = orc.adaptive_default()
    assert (cfg.min_features, cfg.max_features, cfg.max_per_cell, cfg.grid, cfg.edge, cfg.max_iters) == (67, 113, 113, 3, 31, 5)
    th_o = np.full(9, 20.0); th_c = np.full(9, 20.0)
    visited = set()
    for f in range(6):
        img = synth.make_frame(tex, f, w, h, 3)
        ko, found, used = orc.adaptive_detect(img, th_o, retain_best=1000)
        kc, used_c, found_c = cv2_adaptive(cv2, img, th_c, cfg, 1000)
        assert np.array_equal(used, used_c) and np.array_equal(found, found_c) and np.array_equal(th_o, th_c), f"frame {f}"
        assert len(ko) == len(kc)
        got = np.stack([ko["x"], ko["y"], ko["size"], ko["angle"], ko["response"], ko["octave"].astype(np.float32), ko["class_id"].astype(np.float32)], 1)
        assert np.array_equal(got, np.array(kc, np.float32)), f"frame {f}"
        visited |= set(used.tolist())
    assert len(visited) > 2, "the clip must exercise both 'too many' and 'too few' adjustments"


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,n", [(640, 480, 20), (1280, 720, 5)])
def test_cuda_adaptive_route_equals_oracle(ob, orc, w, h, n):
    tex = synth.make_texture(3, h, w)
    frames = np.stack([synth.make_frame(tex, f, w, h, 3) for f in range(n)])
    th_o = np.full(9, 20.0)
    ref = [orc.adaptive_detect(frames[f], th_o, retain_best=1000) for f in range(n)]
    ctx = ob.Context(width=w, height=h, nfeatures=2000 if w > 640 else 1000, max_frames=1)
    try:
        th_g = np.zeros(9)                                   # <= 0 => init_th
        split = n // 2                                       # two calls: the state must carry over (and 20 frames span two sub-batches)
        kps, used, found = ctx.adaptive_detect(frames[:split], th_g)
        k2, u2, f2 = ctx.adaptive_detect(frames[split:], th_g)
        kps += k2; used = np.concatenate([used, u2]); found = np.concatenate([found, f2])
        assert np.array_equal(th_g, th_o)
        for f in range(n):
            assert np.array_equal(used[f], ref[f][2]) and np.array_equal(found[f], ref[f][1]), f"frame {f}: thresholds / counts"
            assert kps[f].tobytes() == ref[f][0].tobytes(), f"frame {f}: keypoints"
        assert ctx.launch_count() >= 4
    finally:
        ctx.close()
