#!/usr/bin/env python
"""Per-stage CUDA-event times of one crf_analyze_crops call.  usage: stage_times.py [faces] [hp_stride] [ffd_stride]"""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import face_alignment_cvpr_2012_b200 as crf
from face_alignment_cvpr_2012_b200 import workloads as wl
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
hs = int(sys.argv[2]) if len(sys.argv) > 2 else 4
fs = int(sys.argv[3]) if len(sys.argv) > 3 else 3
gm = crf.Model(packed=str(wl.staged_model_path()))
crops, _ = wl.make_crops(n)
ctx = crf.Context(gm, 0, crf._options(None, hp_stride=hs, ffd_stride=fs))
ctx.analyze_crops(crops[: min(n, 256)])
ctx.analyze_crops(crops)
ctx.set_profiling(True, False); ctx.reset_counters()
t = time.perf_counter(); ctx.analyze_crops(crops); dt = time.perf_counter() - t
ms, _ = ctx.stage_ms()
print(f"{n} faces strides {hs}/{fs}: wall {dt * 1e3:.2f} ms ({n / dt:.0f} faces/s); stage ms:", {k: round(v, 3) for k, v in ms.items()}, "sum", round(sum(ms.values()), 2))
