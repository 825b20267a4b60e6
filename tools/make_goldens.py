#!/usr/bin/env python3
"""Mints the golden vectors under tests/golden/ (run in the authoring container; cv2 4.13.0 must be importable).

The reference holds no golden vectors, known-answer tests or fixtures for this path, and its OpenCV / PCL
/ Eigen dependencies are absent, so it cannot be run here.  What CAN be pinned is the arithmetic the reference delegates
to OpenCV: every stage below that exists as an OpenCV entry point is produced by calling that entry point through cv2
(the same call the reference makes), NOT by the oracle:
    pyramid levels        cv2.resize(INTER_LINEAR) chained            (orbextractor.cpp:846)
    FAST candidates       cv2.FastFeatureDetector per cell ROI         (orbextractor.cpp:706-712)
    blurred levels        cv2.GaussianBlur(7x7, 2, 2, REFLECT_101)     (orbextractor.cpp:796)
    orientation           cv2.fastAtan2 on integer moments              (orbextractor.cpp:38)
    kNN-2 tables          cv2.BFMatcher(NORM_HAMMING).knnMatch(k=2)     (matcher.cpp:60)
    whole extraction      tests/cv2_oracle.py: the reference's flow re-driven through those cv2 calls
The stages with no library counterpart (quadtree order under the oracle's Q3 tie rule, rBRIEF with no-FMA products, the
sort / sample / Mahalanobis RANSAC) are minted from oracle/ and labelled `oracle_defined` in the manifest.

  python tools/make_goldens.py          # rewrites tests/golden/*.npz + manifest.json
"""
import hashlib
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "tests")]
import cv2  # noqa: E402
import cv2_oracle as co  # noqa: E402
import synth  # noqa: E402
from oracle import oracle as orc  # noqa: E402

OUT = ROOT / "tests" / "golden"
OUT.mkdir(exist_ok=True)
manifest = {"cv2": cv2.__version__, "numpy": np.__version__, "files": {}}


def kp_array(kps):
    a = np.zeros(len(kps), orc.KEYPOINT_DT)
    for i, k in enumerate(kps):
        a[i] = k
    return a


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def save(name, note, **arrays):
    np.savez_compressed(OUT / name, **arrays)
    manifest["files"][name] = {"note": note, "arrays": {k: [str(v.dtype), list(v.shape)] for k, v in arrays.items()},
                               "bytes": (OUT / name).stat().st_size}


def extraction_case(name, w, h, nfeatures, nlevels, frame_ids, seed):
    """Small frames: full per-stage goldens.  Inputs are regenerated from tests/synth.py seeds; their sha256 is stored."""
    orc.build()
    tex = synth.make_texture(seed, h, w)
    pattern = orc.pattern()
    arrays = {"params": np.array([w, h, nfeatures, nlevels, seed], np.int32), "frame_ids": np.array(frame_ids, np.int32)}
    for i in frame_ids:
        img = synth.make_frame(tex, i, w, h, seed)
        depth = synth.make_depth(i, w, h, seed)
        kps, desc, dbg = co.extract(img, pattern, nfeatures=nfeatures, nlevels=nlevels)            # cv2-driven flow
        kps = kp_array(kps)
        levels = co.pyramid(img, 1.2, nlevels)
        blurred = [cv2.GaussianBlur(l.copy(), (7, 7), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101) for l in levels]
        cands = [np.array(co.fast_cells(l), np.int32).reshape(-1, 3) for l in levels]
        ko, do = orc.extract(img, nfeatures=nfeatures, nlevels=nlevels)
        # the cv2-driven flow and the C oracle agree bit for bit before anything is frozen
        assert kps.tobytes() == ko.tobytes() and np.array_equal(desc, do), "cv2 flow and oracle disagree: fix before minting"
        xyz, uright = orc.unproject(ko, depth)
        p = f"f{i}_"
        arrays[p + "input_sha"] = np.frombuffer(bytes.fromhex(sha(img)), np.uint8)
        arrays[p + "depth_sha"] = np.frombuffer(bytes.fromhex(sha(depth)), np.uint8)
        arrays[p + "pyramid"] = np.concatenate([l.ravel() for l in levels])
        arrays[p + "blurred"] = np.concatenate([b.ravel() for b in blurred])
        arrays[p + "level_wh"] = np.array([[l.shape[1], l.shape[0]] for l in levels], np.int32)
        arrays[p + "cand_counts"] = np.array([len(c) for c in cands], np.int32)
        arrays[p + "cands"] = np.concatenate(cands) if cands else np.zeros((0, 3), np.int32)
        arrays[p + "kp_counts"] = np.array([len(k) for k in dbg["kept"]], np.int32)
        arrays[p + "keypoints"] = kps
        arrays[p + "descriptors"] = desc
        arrays[p + "xyz"] = xyz
        arrays[p + "uright"] = uright
    save(name, "per-stage extraction goldens: pyramid / blur / FAST candidates / angles via cv2 entry points; quadtree order, "
         "descriptors and unprojection oracle_defined (and equal to the cv2-driven flow of tests/cv2_oracle.py)", **arrays)


def full_size_case(name, frame_ids, seed):
    """640x480, 1000 kp, 8 levels (BASELINE configs[0]): outputs + stage checksums only, to keep the fixture small."""
    tex = synth.make_texture(seed, 480, 640)
    pattern = orc.pattern()
    arrays = {"frame_ids": np.array(frame_ids, np.int32)}
    for i in frame_ids:
        img = synth.make_frame(tex, i, 640, 480, seed)
        kps, desc, dbg = co.extract(img, pattern)
        kps = kp_array(kps)
        ko, do = orc.extract(img)
        assert kps.tobytes() == ko.tobytes() and np.array_equal(desc, do)
        levels = co.pyramid(img)
        blurred = [cv2.GaussianBlur(l.copy(), (7, 7), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101) for l in levels]
        p = f"f{i}_"
        arrays[p + "input_sha"] = np.frombuffer(bytes.fromhex(sha(img)), np.uint8)
        arrays[p + "pyramid_sha"] = np.stack([np.frombuffer(bytes.fromhex(sha(l)), np.uint8) for l in levels])
        arrays[p + "blurred_sha"] = np.stack([np.frombuffer(bytes.fromhex(sha(b)), np.uint8) for b in blurred])
        arrays[p + "cand_counts"] = np.array([len(co.fast_cells(l)) for l in levels], np.int32)
        arrays[p + "kp_counts"] = np.array([len(k) for k in dbg["kept"]], np.int32)
        arrays[p + "keypoints"] = kps
        arrays[p + "descriptors"] = desc
    save(name, "640x480 / 1000 kp / 8 levels: keypoints + descriptors of the cv2-driven flow (== oracle), sha256 of every pyramid "
         "and blurred level from cv2.resize / cv2.GaussianBlur", **arrays)


def matching_case(name):
    bf = cv2.BFMatcher(cv2.NORM_HAMMING)
    arrays = {}
    for tag, (A, B) in {"rand": synth.descriptor_sets(1000, seed=1), "ties": synth.tie_heavy_sets(1000, seed=3),
                        "ragged": tuple(x[:n] for x, n in zip(synth.descriptor_sets(700, seed=9), (613, 257))),
                        "tiny": tuple(x[:n] for x, n in zip(synth.descriptor_sets(64, seed=11), (5, 2)))}.items():
        knn = bf.knnMatch(A, B, k=2)
        t = np.array([[m[0].trainIdx, int(m[0].distance), m[1].trainIdx, int(m[1].distance)] for m in knn], np.int32)
        arrays[tag + "_A_sha"] = np.frombuffer(bytes.fromhex(sha(A)), np.uint8)
        arrays[tag + "_B_sha"] = np.frombuffer(bytes.fromhex(sha(B)), np.uint8)
        arrays[tag + "_knn"] = t
        for ratio in (0.6, 0.8, 0.9):       # Matcher::KnnMatch's ratio test on cv2's float distances (matcher.cpp:64)
            keep = [i for i, m in enumerate(knn) if np.float32(m[0].distance) < np.float32(ratio) * np.float32(m[1].distance)]
            arrays[f"{tag}_ratio{int(ratio * 10)}"] = np.array(keep, np.int32)
        arrays[tag + "_cross8"] = orc.knn_match(A, B, 0.8, True)["queryIdx"].astype(np.int32)     # oracle_defined (quirk Q10)
    save(name, "kNN-2 tables and ratio survivors from cv2.BFMatcher(NORM_HAMMING).knnMatch(k=2); *_cross8 (mutual-NN, a north-star "
         "extension with no reference counterpart) oracle_defined", **arrays)


def ransac_case(name):
    arrays = {}
    for tag, seed, m in (("a", 4, 650), ("b", 5, 300), ("few", 6, 24)):
        src, dst, matches, R, t = synth.rigid_pairs(m=m, seed=seed)
        r = orc.ransac_iterate(src, dst, matches, seed=42)
        arrays[tag + "_good_sorted"] = r["good_sorted"]
        arrays[tag + "_sample_table"] = r["sample_table"]
        arrays[tag + "_inliers"] = r["inliers"]
        arrays[tag + "_T12"] = r["T12"]
        arrays[tag + "_scalars"] = np.array([r["ok"], r["n_good"], r["real_iters"], r["valid_iters"], r["used_identity"]], np.int32)
        arrays[tag + "_rmse_cov"] = np.array([r["rmse"], r["depth_cov"]], np.float64)
        arrays[tag + "_hyp_n"] = r["hyp"]["n_refined"].astype(np.int32)
        arrays[tag + "_truth_Rt"] = np.concatenate([R.ravel(), t]).astype(np.float64)
    rng = np.random.default_rng(8)
    A = rng.normal(size=(40, 3)).astype(np.float32)
    ang = 0.3
    Rz = np.array([[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]], np.float32)
    B = (A @ Rz.T + np.array([0.1, -0.2, 0.3], np.float32)).astype(np.float32)
    arrays["kabsch_T"] = orc.kabsch(A, B)
    arrays["kabsch_reflect_T"] = orc.kabsch(A, (A * np.array([1, 1, -1], np.float32)).astype(np.float32))
    arrays["kabsch_empty_T"] = orc.kabsch(np.zeros((0, 3), np.float32), np.zeros((0, 3), np.float32))
    arrays["libc_rand_seed42"] = orc.libc_rand_sequence(42, 64)
    save(name, "RANSAC(200,20,3.0,4) on seeded rigid pairs: std::sort order, glibc-rand sample table, inlier set, pose — oracle_defined "
         "(PCL / Eigen absent; libc rand() and libstdc++ std::sort are the real library calls inside the oracle)", **arrays)


if __name__ == "__main__":
    orc.build()
    extraction_case("extract_320x240.npz", 320, 240, 300, 6, [0, 7], seed=2)
    full_size_case("extract_640x480.npz", [0, 7, 13], seed=0)
    matching_case("match_knn2.npz")
    ransac_case("ransac.npz")
    (OUT / "manifest.json").write_text(json.dumps(manifest, indent=1) + "\n")
    print(json.dumps({k: v["bytes"] for k, v in manifest["files"].items()}))
