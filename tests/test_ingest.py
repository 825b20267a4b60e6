"""Frame ingest, the step right before the path: Frame::Frame converts the colour image with
cv::cvtColor(CV_BGR2GRAY) (Core/frame.cpp:23) before ExtractFeatures.  The conversion runs on the device in
orbf_extract_batch_bgr; bit-exact against OpenCV's fixed-point arithmetic."""
import numpy as np
import pytest

import synth


def colour_frames(texture, ids, w=640, h=480):
    """Synthetic BGR frames: three differently shifted / scaled views of the gray synthetic frame, so channels differ."""
    out = []
    for i in ids:
        g = synth.make_frame(texture, i, w, h).astype(np.int32)
        b = np.clip(np.roll(g, 3, axis=1) * 9 // 10 + 7, 0, 255); r = np.clip(np.roll(g, -2, axis=0) * 11 // 10 - 5, 0, 255)
        out.append(np.stack([b, g, r], axis=-1).astype(np.uint8))
    return np.ascontiguousarray(np.stack(out))


def test_oracle_bgr2gray_equals_cv2(orc, texture):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(0)
    for img in (rng.integers(0, 256, (480, 640, 3), dtype=np.uint8), rng.integers(0, 256, (7, 13, 3), dtype=np.uint8),
                colour_frames(texture, [4])[0], np.full((5, 5, 3), 255, np.uint8), np.zeros((5, 5, 3), np.uint8)):
        assert np.array_equal(orc.bgr2gray(img), cv2.cvtColor(img, cv2.COLOR_BGR2GRAY))


@pytest.mark.gpu
def test_cuda_bgr_ingest_equals_oracle(ob, orc, texture):
    bgr = colour_frames(texture, [0, 7, 13])
    depths = np.stack([synth.make_depth(i) for i in (0, 7, 13)])
    ctx = ob.Context(max_frames=3)
    try:
        ctx.extract_batch_bgr(bgr, depths)
        for s in range(3):
            gray = orc.bgr2gray(bgr[s])
            assert np.array_equal(ctx.download_gray(s), gray), f"gray plane {s}"
            k, d, xyz = ctx.download_frame(s)
            ko, do = orc.extract(gray)
            assert k.tobytes() == ko.tobytes() and np.array_equal(d, do) and np.array_equal(xyz, orc.unproject(ko, depths[s])[0])
    finally:
        ctx.close()


@pytest.mark.gpu
def test_cuda_bgr_ingest_odd_width(ob, orc):
    """Width not a multiple of 4 and a padded row stride exercise the scalar edge path of the conversion kernel."""
    rng = np.random.default_rng(5)
    w, h = 322, 242
    bgr = rng.integers(0, 256, (2, h, w, 3), dtype=np.uint8)
    ctx = ob.Context(width=w, height=h, nfeatures=300, nlevels=6, max_frames=2)
    try:
        ctx.extract_batch_bgr(bgr)
        for s in range(2):
            assert np.array_equal(ctx.download_gray(s), orc.bgr2gray(bgr[s]))
    finally:
        ctx.close()
