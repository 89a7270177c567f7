#!/usr/bin/env python
"""A/B of the two shared-memory-window traversals: k_traverse_win (CRF_WIN_FMT=1) against k_traverse_win2 (CRF_WIN_FMT=2, internal-nodes-only
records, software-pipelined walks), over launch shapes (warps per CTA | walks per lane).  usage: win2_variants.py [faces=2048]"""
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
n = sys.argv[1] if len(sys.argv) > 1 else "2048"
variants = [(1, "30|1", "20|2"), (2, "30|1", "20|2"), (2, "15|2", "24|2"), (2, "32|1", "20|1"), (2, "20|1", "15|2")]
for fmt, hp, ffd in variants:
    env = dict(os.environ)
    f = lambda s: str(int(s.split("|")[0]) | int(s.split("|")[1]) << 8)
    env["CRF_WIN_FMT"], env["CRF_WIN_HP"], env["CRF_WIN_FFD"] = str(fmt), f(hp), f(ffd)
    r = subprocess.run([sys.executable, str(ROOT / "tools" / "stage_times.py"), n, "1", "1"], env=env, capture_output=True, text=True)
    line = (r.stdout.strip().splitlines() or [r.stderr[-300:]])[-1]
    i = line.find("'hp_traverse'")
    print(f"fmt {fmt} hp {hp:5s} ffd {ffd:5s}:", line[i:i + 80] if i >= 0 else line[-300:], flush=True)
