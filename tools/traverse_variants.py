#!/usr/bin/env python
"""Times the forest-traversal kernel variants (CRF_TRAVERSE_VARIANT = LW | MODE << 8 | NW << 16) on the C2 workload."""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import face_alignment_cvpr_2012_b200 as crf  # noqa: E402
from face_alignment_cvpr_2012_b200 import workloads as wl  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
hs = int(sys.argv[2]) if len(sys.argv) > 2 else 1
fs = int(sys.argv[3]) if len(sys.argv) > 3 else 1
gm = crf.Model(packed=str(wl.staged_model_path()))
crops, _ = wl.make_crops(n)
ref = None
variants = [(10, 32, 2), (10, 32, 4), (15, 32, 4)]
for nw, lw, mode in variants:
    os.environ["CRF_TRAVERSE_VARIANT"] = str(lw | (mode << 8) | (nw << 16))
    ctx = crf.Context(gm, 0, crf._options(None, hp_stride=hs, ffd_stride=fs))
    ctx.analyze_crops(crops[:64])
    ctx.set_profiling(True, False)
    ctx.reset_counters()
    for _ in range(2):
        out = ctx.analyze_crops(crops)
    ms, _ = ctx.stage_ms()
    same = True if ref is None else out.tobytes() == ref.tobytes()
    if ref is None:
        ref = out
    print(f"NW={nw:2d} LW={lw:2d} MODE={mode}: hp_traverse {ms['hp_traverse'] / 2:8.3f} ms  ffd_traverse {ms['ffd_traverse'] / 2:8.3f} ms  identical={same}", flush=True)
    ctx.close()
