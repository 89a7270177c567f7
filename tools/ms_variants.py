#!/usr/bin/env python
"""Times k_meanshift register/occupancy variants (CRF_MS_VARIANT) and chunk sizes on the C2 workload."""
import os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import face_alignment_cvpr_2012_b200 as crf
from face_alignment_cvpr_2012_b200 import workloads as wl
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
gm = crf.Model(packed=str(wl.staged_model_path()))
crops, _ = wl.make_crops(n)
for v in (2, 3, 4):
    for chunk in (1024, 2048):
        os.environ["CRF_MS_VARIANT"] = str(v)
        ctx = crf.Context(gm, 0, crf._options(None, hp_stride=1, ffd_stride=1, max_chunk=chunk))
        ctx.analyze_crops(crops[:64]); ctx.set_profiling(True, False); ctx.reset_counters()
        ctx.analyze_crops(crops)
        ms, _ = ctx.stage_ms()
        print(f"MINB={v} chunk={chunk}: meanshift {ms['meanshift']:.2f} ms votes {ms['votes']:.2f} hp_reduce {ms['hp_reduce']:.2f} (for {n} faces)", flush=True)
        ctx.close()
