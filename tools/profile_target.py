#!/usr/bin/env python
"""Small fixed workload for ncu (one process, one GPU): N crops of the C2 workload through the host C-ABI call.
usage: profile_target.py [faces=512] [hp_stride=1] [ffd_stride=1] [steps=2]"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import face_alignment_cvpr_2012_b200 as crf  # noqa: E402
from face_alignment_cvpr_2012_b200 import workloads as wl  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
hs = int(sys.argv[2]) if len(sys.argv) > 2 else 1
fs = int(sys.argv[3]) if len(sys.argv) > 3 else 1
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
gm = crf.Model(packed=str(wl.staged_model_path()))
crops, _ = wl.make_crops(n)
ctx = crf.Context(gm, 0, crf._options(None, hp_stride=hs, ffd_stride=fs))
for _ in range(steps):
    out = ctx.analyze_crops(crops)
print("ok", n, float(out["headpose"][0]), ctx.counters()["kernel_launches"])
