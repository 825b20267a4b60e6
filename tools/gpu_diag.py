"""Stage-by-stage GPU-vs-oracle mismatch report (authoring aid; run on the GPU box)."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "tests")]
from conftest import load_orbfront  # noqa: E402
import synth  # noqa: E402
from oracle import oracle as orc  # noqa: E402

ob = load_orbfront()
orc.build()
tex = synth.make_texture(0, 480, 640)
img = synth.make_frame(tex, 0)
depth = synth.make_depth(0)
ctx = ob.Context(max_frames=2)
ctx.extract_batch(img[None], depth[None])
kps_o, desc_o, dbg = orc.extract(img, debug=True)
levels_o = orc.split_levels(dbg["pyramid"], dbg["ws"], dbg["hs"])
blur_o = orc.split_levels(dbg["blurred"], dbg["ws"], dbg["hs"])
off = 0
print("level kp counts gpu", ctx.level_keypoint_counts(0), "oracle", dbg["n_kps"])
for l in range(8):
    p = ctx.pyramid_level(0, l)
    b = ctx.pyramid_level(0, l, True)
    c = ctx.level_candidates(0, l)
    n = int(dbg["n_cands"][l]); ref = dbg["cands"][off:off + n]; off += n
    same = len(c) == n and np.array_equal(c, ref)
    print(f"L{l}: pyr diff {(p != levels_o[l]).sum()}  blur diff {(b != blur_o[l]).sum()}  cands {len(c)} vs {n} equal={same}")
    if not same and len(c) and n:
        k = min(len(c), n)
        bad = np.nonzero((c[:k]["x"] != ref[:k]["x"]) | (c[:k]["y"] != ref[:k]["y"]) | (c[:k]["score"] != ref[:k]["score"]))[0]
        print("   first diffs at", bad[:5], c[bad[:3]], ref[bad[:3]])
        sg = set(map(tuple, c.tolist())); sr = set(map(tuple, ref.tolist()))
        print("   set equal:", sg == sr, "only gpu", list(sg - sr)[:5], "only ref", list(sr - sg)[:5])
kps, desc, xyz = ctx.download_frame(0)
print("kps", len(kps), len(kps_o))
n = min(len(kps), len(kps_o))
for f in kps.dtype.names:
    bad = np.nonzero(kps[f][:n] != kps_o[f][:n])[0]
    print(f"  field {f}: {len(bad)} diffs", bad[:5], kps[f][bad[:3]], kps_o[f][bad[:3]])
print("desc rows differing:", (desc[:n] != desc_o[:n]).any(axis=1).sum())
xyz_o, _ = orc.unproject(kps_o, depth)
print("xyz rows differing:", (xyz[:n] != xyz_o[:n]).any(axis=1).sum())
