#!/usr/bin/env python
"""Per-source-line instruction / stall-sample breakdown of one kernel from an ncu report (source page), plus a few headline metrics.
usage: ncu_lines.py report.ncu-rep [unit_count] [min_pct]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0; minpct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.8
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
want = ["gpu__time_duration.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__registers_per_thread"]
for a, b in zip(rows[0], rows[2]):
    if a in want or (a.startswith("smsp__average_warps_issue_stalled") and a.endswith("per_issue_active.ratio") and float(b or 0) > 0.15): print(f"{a:90s} {b}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = next(r for r in rows if "Line No" in r)
iL, iI, iS, iSrc = hdr.index("Line No"), hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
agg = {}
for r in rows:
    if len(r) != len(hdr): continue
    try: ln = int(r[iL]); ins = float(r[iI] or 0); smp = float(r[iS] or 0)
    except ValueError: continue
    a = agg.setdefault(ln, [0, 0, r[iSrc]]); a[0] += ins; a[1] += smp
tot = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
print(f"instructions {tot:.0f} ({tot / units:.1f} per unit), samples {ts:.0f}")
for ln, (i, s, text) in sorted(agg.items()):
    if i / tot * 100 > minpct or s / ts * 100 > minpct: print(f"{ln:4d} {i / units:8.1f}/unit {i / tot * 100:5.1f}% inst {s / ts * 100:5.1f}% samp  {text[:96]}")
