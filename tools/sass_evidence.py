#!/usr/bin/env python3
"""Writes profiles/sass_evidence.md: per-kernel counts of the SASS mnemonics that show how each kernel is built
(UTMALDG = TMA tile loads, VIMNMX3.U16x2 / VABSDIFF4 / IDP.4A / IDP.2A = packed integer SIMD, LOP3 + POPC = matcher)."""
import collections
import re
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
so = ROOT / "adaptive-rgbd-localization-mappig_b200" / "liborbfront_b200.so"
txt = subprocess.run(["cuobjdump", "-sass", str(so)], capture_output=True, text=True).stdout
kern = None
counts = collections.defaultdict(collections.Counter)
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(anonymous namespace\)::", "", kern).split("(")[0]
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Za-z0-9_.]+)", line)
    if m and kern:
        counts[kern][m.group(1)] += 1
want = ["UTCIMMA", "LDTM.x32", "UTCBAR", "UTMALDG.3D", "SYNCS.ARRIVE.TRANS64", "SYNCS.PHASECHK.TRANS64.TRYWAIT", "VIMNMX3.U16x2", "VIMNMX.U16x2", "VABSDIFF4.U8", "IDP.4A", "IDP.2A",
        "POPC", "LOP3.LUT", "CREDUX.MIN", "ATOMS", "DFMA", "SHFL", "VOTE", "BAR.SYNC"]
lines = ["# SASS evidence (cuobjdump -sass liborbfront_b200.so, sm_100a): instruction counts per kernel", "",
         "TMA = `UTMALDG.3D` (+ mbarrier `SYNCS.*`); packed 16x2 integer min/max = `VIMNMX(3).U16x2`; byte SIMD = `VABSDIFF4`, `IDP.4A`, "
         "`IDP.2A`; matcher = `UTCIMMA` (tcgen05.mma kind::i8), `LDTM` (tcgen05.ld), `UTCBAR` (tcgen05.commit) + packed `VIMNMX.U16x2` epilogue.", "Regenerate with `python tools/sass_evidence.py`.", "",
         "| kernel | total | " + " | ".join(want) + " |", "|---|---:|" + "---:|" * len(want)]
for k in sorted(counts, key=lambda k: -sum(counts[k].values())):
    c = counts[k]
    row = [str(sum(c.values()))]
    for w in want:
        n = sum(v for op, v in c.items() if op == w or op.startswith(w + "."))
        row.append(str(n) if n else "")
    lines.append(f"| `{k}` | " + " | ".join(row) + " |")
(ROOT / "profiles" / "sass_evidence.md").write_text("\n".join(lines) + "\n")
print("\n".join(lines))
