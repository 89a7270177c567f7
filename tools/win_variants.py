#!/usr/bin/env python
"""Sweep of the k_traverse_win launch shapes (warps per CTA | walks per lane << 8) through CRF_WIN_HP / CRF_WIN_FFD.
usage: win_variants.py [faces=2048]   (development aid; prints the traversal stage times of each variant)"""
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
n = sys.argv[1] if len(sys.argv) > 1 else "2048"
variants = [("30|1", "20|2"), ("15|2", "20|2"), ("32|1", "20|1"), ("16|1", "24|1"), ("20|1", "28|1"), ("24|1", "32|1"), ("30|1", "10|2")]
if os.environ.get("WIN_SWEEP") == "pairx":   # p = two walks per lane on horizontal neighbours with the shared record fetch
    variants = [("30|1", "20|2"), ("15|2p", "20|2p"), ("16|2p", "24|2p"), ("20|2p", "10|2p"), ("24|2p", "15|2p")]
for hp, ffd in variants:
    env = dict(os.environ)
    f = lambda s: str(int(s.split("|")[0]) | int(s.split("|")[1].rstrip("p")) << 8 | (1 << 12 if s.endswith("p") else 0))
    env["CRF_WIN_HP"], env["CRF_WIN_FFD"] = f(hp), f(ffd)
    r = subprocess.run([sys.executable, str(ROOT / "tools" / "stage_times.py"), n, "1", "1"], env=env, capture_output=True, text=True)
    line = (r.stdout.strip().splitlines() or [r.stderr[-300:]])[-1]
    i = line.find("'hp_traverse'")
    print(f"hp {hp:5s} ffd {ffd:5s}:", line[i:i + 80] if i >= 0 else line[-200:], flush=True)
