#!/usr/bin/env python
"""Extracts per-launch DRAM traffic of the traversal kernels from an `ncu --set full` report into profiles/r1_traffic.json.
usage: ncu_traffic.py <report.ncu-rep> <faces_per_launch>"""
import csv
import json
import subprocess
import sys
from pathlib import Path

rep, faces = sys.argv[1], int(sys.argv[2])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}


def to_bytes(v, u):
    return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u]


trav = [r for r in rows[2:] if "k_traverse" in r[ix["Kernel Name"]]]
out = {}
for name, r in zip(["k_traverse_hp", "k_traverse_ffd"], trav[:2]):   # launch order: head pose, then FFD
    rd = to_bytes(r[ix["dram__bytes_read.sum"]], units[ix["dram__bytes_read.sum"]])
    wr = to_bytes(r[ix["dram__bytes_write.sum"]], units[ix["dram__bytes_write.sum"]])
    out[name] = {"dram_bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr, "faces_per_launch": faces,
                 "duration_ms_under_ncu": float(r[ix["gpu__time_duration.sum"]]), "report": Path(rep).name}
Path(__file__).resolve().parents[1].joinpath("profiles", "r1_traffic.json").write_text(json.dumps(out, indent=1))
print(json.dumps(out, indent=1))
