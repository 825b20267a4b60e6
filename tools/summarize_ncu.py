#!/usr/bin/env python3
"""Summarise ncu outputs brought back in gpurun_out/ into profiles/ (tracked).
  python tools/summarize_ncu.py <tag> [launches.csv] [raw.csv]"""
import collections
import csv
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
tag = sys.argv[1]
launches = Path(sys.argv[2]) if len(sys.argv) > 2 else ROOT / "gpurun_out" / f"launches_{tag}.csv"
raw = Path(sys.argv[3]) if len(sys.argv) > 3 else ROOT / "gpurun_out" / f"prof_{tag}_raw.csv"
out = ROOT / "profiles"
out.mkdir(exist_ok=True)

if launches.exists():
    rows = [r for r in csv.reader(open(launches)) if len(r) > 5]
    start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[start]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[start + 1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        v_us = v / 1e3 if r[ui] in ("ns", "nsecond") else (v if r[ui] in ("us", "usecond") else v * 1e3)
        n = r[ki].split("(")[0].replace("<unnamed>::", "").replace("void ", "")
        a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += v_us
    tot = sum(v[1] for v in agg.values())
    lines = [f"# ncu launch list `{tag}` (gpu__time_duration.sum, --clock-control none; cold-cache, serialised: compare SHARES)", "",
             "| kernel | launches | total us | share |", "|---|---:|---:|---:|"]
    for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        lines.append(f"| {n} | {c} | {t:.1f} | {100 * t / tot:.1f}% |")
    (out / f"{tag}_launches.md").write_text("\n".join(lines) + "\n")
    (out / f"{tag}_launches.csv").write_text(launches.read_text())
    print("\n".join(lines))

if raw.exists():
    rows = list(csv.reader(open(raw)))
    hdr = rows[0]
    want = ["Kernel Name", "launch__grid_size", "launch__block_size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "l1tex__throughput.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "smsp__inst_executed.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum",
            "sm__inst_executed_pipe_lsu.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__occupancy_limit_registers",
            "launch__occupancy_limit_shared_mem"]
    idx = [(w, hdr.index(w)) for w in want if w in hdr]
    with open(out / f"{tag}_kernels.csv", "w", newline="") as f:
        wr = csv.writer(f)
        wr.writerow([w for w, _ in idx]); wr.writerow([rows[1][j] for _, j in idx])
        for r in rows[2:]:
            wr.writerow([r[j].split("(")[0] if w == "Kernel Name" else r[j] for w, j in idx])
    print("wrote", out / f"{tag}_kernels.csv", len(rows) - 2, "kernel launches")
