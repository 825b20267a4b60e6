#!/usr/bin/env python
"""Per-phase instruction split of one kernel from an ncu report with source (ncu --set full --import-source on).
usage: ncu_phases.py report.ncu-rep units "phase=file.cu:lo-hi[,file.h:lo-hi...]" ...   (units = e.g. pixels processed by the launch)
Lines are keyed by (file, line) — fast.cu and fast_device.h share line numbers.  Prints, per phase: warp instructions, thread
instructions (lanes that executed), both per unit, share of the kernel's warp instructions and of its stall samples."""
import csv, io, subprocess, sys
rep = sys.argv[1]; units = float(sys.argv[2])
phases = []
for spec in sys.argv[3:]:
    name, rng = spec.split("=")
    parts = []
    for p in rng.split(","):
        f, lr = p.split(":"); lo, hi = lr.split("-")
        parts.append((f, int(lo), int(hi)))
    phases.append((name, parts))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur = None; hdr = None; agg = {}
for r in csv.reader(io.StringIO(src)):
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]; continue
    if "Line No" in r:
        hdr = r; iL, iI, iT, iS = hdr.index("Line No"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples"); continue
    if hdr is None or len(r) != len(hdr) or not r[iL]:
        continue
    try:
        ln = int(r[iL]); a = agg.setdefault((cur, ln), [0.0, 0.0, 0.0, r[1]])
        a[0] += float(r[iI] or 0); a[1] += float(r[iT] or 0); a[2] += float(r[iS] or 0)
    except ValueError:
        pass
tw = sum(v[0] for v in agg.values()); tt = sum(v[1] for v in agg.values()); ts = sum(v[2] for v in agg.values())
print(f"kernel: {tw:.0f} warp instr ({tw / units:.3f}/unit), {tt:.0f} thread instr ({tt / units:.2f}/unit, {tt / tw:.1f} lanes/instr), {ts:.0f} samples")
left = dict(agg)
print(f"{'phase':28s} {'warp/unit':>10s} {'thread/unit':>12s} {'% warp':>7s} {'% samp':>7s}")
for name, parts in phases:
    w = t = s = 0.0
    for (f, ln), v in list(left.items()):
        if any(f == pf and lo <= ln <= hi for pf, lo, hi in parts):
            w += v[0]; t += v[1]; s += v[2]; del left[(f, ln)]
    print(f"{name:28s} {w / units:10.3f} {t / units:12.2f} {w / tw * 100:7.1f} {s / ts * 100:7.1f}")
w = sum(v[0] for v in left.values()); t = sum(v[1] for v in left.values()); s = sum(v[2] for v in left.values())
print(f"{'(other)':28s} {w / units:10.3f} {t / units:12.2f} {w / tw * 100:7.1f} {s / ts * 100:7.1f}")
for (f, ln), v in sorted(left.items(), key=lambda kv: -kv[1][0])[:8]:
    print(f"    other: {f}:{ln} {v[0] / tw * 100:.1f}%  {v[3][:80]}")
