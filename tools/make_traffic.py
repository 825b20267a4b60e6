#!/usr/bin/env python3
"""profiles/traffic.json from profiles/<tag>_kernels.csv (one ncu --set full capture per kernel of a 128-frame step):
dram__bytes_read.sum + dram__bytes_write.sum per stage, per step and per frame.   usage: make_traffic.py <tag> [frames]"""
import csv, json, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
tag = sys.argv[1]; frames = int(sys.argv[2]) if len(sys.argv) > 2 else 128
rows = list(csv.reader(open(ROOT / "profiles" / f"{tag}_kernels.csv")))
h = rows[0]; units = rows[1]
ik, it, ir, iw = h.index("Kernel Name"), h.index("gpu__time_duration.sum"), h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum")
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
stage_of = {"resize_tile_kernel": "pyr_resize", "fast_strip_kernel": "fast_cell", "quadtree_kernel": "quadtree", "blur_tile_kernel": "blur7", "describe_kernel": "describe",
            "knn2_kernel": "hamming_knn2"}
out = {"source": f"profiles/{tag}_kernels.csv (ncu --set full, bench.py --frames {frames}): dram__bytes_read.sum + dram__bytes_write.sum summed over the stage's "
                 f"launches of one step, bytes per step of {frames} frames", "frames_per_step": frames}
for r in rows[2:]:
    name = r[ik].replace("<unnamed>::", "").replace("void ", "").split("<")[0].split("(")[0]
    st = stage_of.get(name)
    if not st:
        continue
    b = float(r[ir]) * scale[units[ir]] + float(r[iw]) * scale[units[iw]]
    e = out.setdefault(st, {"dram_bytes_per_step": 0.0, "launches": 0, "ncu_us_per_step": 0.0})
    e["dram_bytes_per_step"] += b; e["launches"] += 1; e["ncu_us_per_step"] += float(r[it])
for st, e in out.items():
    if isinstance(e, dict):
        e["dram_bytes_per_frame"] = e["dram_bytes_per_step"] / frames
(ROOT / "profiles" / "traffic.json").write_text(json.dumps(out, indent=1) + "\n")
print(json.dumps({k: (round(v["dram_bytes_per_frame"]), v["launches"]) for k, v in out.items() if isinstance(v, dict)}))
