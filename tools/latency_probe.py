#!/usr/bin/env python
"""Stage times of single-face / small-batch calls at the reference's default strides (development aid)."""
import sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import face_alignment_cvpr_2012_b200 as crf
from face_alignment_cvpr_2012_b200 import workloads as wl
gm = crf.Model(packed=str(wl.staged_model_path()))
crops, _ = wl.make_crops(64)
ctx = crf.Context(gm, 0)
for n in (1, 16, 64):
    for _ in range(5):
        ctx.analyze_crops(crops[:n])
    ctx.set_profiling(True, False); ctx.reset_counters()
    reps = 20
    t = time.perf_counter()
    for _ in range(reps):
        ctx.analyze_crops(crops[:n])
    dt = (time.perf_counter() - t) / reps
    ms, _ = ctx.stage_ms()
    print(f"n={n}: wall {dt * 1e3:.3f} ms/call; stage ms/call:", {k: round(v / reps, 4) for k, v in ms.items()}, "sum", round(sum(ms.values()) / reps, 3), flush=True)
    ctx.set_profiling(False, False)
    t = time.perf_counter()
    for _ in range(reps):
        ctx.analyze_crops(crops[:n])
    print(f"      no profiling: {(time.perf_counter() - t) / reps * 1e3:.3f} ms/call")
