"""A short run of the randomised soak (tools/soak_oracle_vs_ref.py) inside the CPU suite: the oracle against the reference's OWN sources
(oracle/_ref) and against the real cv2 on problems drawn at random — geometry, parameters and content are not the fixed cases of the
other tests.  All modes run side by side for a few seconds each; the long runs are recorded in profiles/r3t_soak_oracle_vs_ref.jsonl.

CPU-only.  Skipped where neither the reference checkout nor a prebuilt oracle/_ref exists."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
MODES = ["extract", "ransac", "match", "adaptive", "8f", "frame", "landmark", "odometry", "cv2", "replay", "linalg"]


def test_short_soak_finds_no_disagreement(orc):
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built and /root/reference absent")
    procs = [subprocess.Popen([sys.executable, "-W", "ignore", str(ROOT / "tools" / "soak_oracle_vs_ref.py"), m, "20261019", "5"],
                              stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True) for m in MODES]
    for m, p in zip(MODES, procs):
        out, _ = p.communicate(timeout=300)
        assert p.returncode == 0, m
        r = json.loads(out.strip().splitlines()[-1])
        assert r["mode"] == m and r["cases"] >= 1, r
        assert r["mismatches"] == 0, r
