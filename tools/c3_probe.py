#!/usr/bin/env python
"""C3 (1080p frames x 16 faces, default strides, host frames): wall time of one crf_analyze_batch call against the kernel stages,
for the library's chunking and for explicit max_chunk values.  usage: c3_probe.py [frames=64]"""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np
import torch
import face_alignment_cvpr_2012_b200 as crf
from face_alignment_cvpr_2012_b200 import workloads as wl
nf = int(sys.argv[1]) if len(sys.argv) > 1 else 64
gm = crf.Model(packed=str(wl.staged_model_path()))
frames, boxes, iob, _ = wl.make_frames(nf, seed=2013)
h = torch.from_numpy(frames).pin_memory()
fr = h.numpy()
for chunk in (0, 1024, 512):
    ctx = crf.Context(gm, 0, crf._options(None, max_chunk=chunk))
    for _ in range(2):
        ctx.analyze_batch(fr, boxes, iob)
    ctx.set_profiling(True, False); ctx.reset_counters()
    t = time.perf_counter(); ctx.analyze_batch(fr, boxes, iob); dt = time.perf_counter() - t
    ms, _ = ctx.stage_ms()
    ctx.set_profiling(False, False)
    t = time.perf_counter()
    for _ in range(3):
        ctx.analyze_batch(fr, boxes, iob)
    dt2 = (time.perf_counter() - t) / 3
    c = ctx.counters()
    print(f"max_chunk {chunk}: {len(boxes)} faces, wall {dt2 * 1e3:.2f} ms ({len(boxes) / dt2:.0f} faces/s); kernels {sum(ms.values()):.2f} ms", {k: round(v, 2) for k, v in ms.items()}, "h2d MB", c["h2d_bytes"] / 1e6 / 6, flush=True)
    ctx.close()
