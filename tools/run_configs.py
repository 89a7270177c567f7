#!/usr/bin/env python
"""Throughput of the other BASELINE.json configurations on one GPU (they are parity-test cases, not bench lines; this is
the measured context for DESIGN.md).  Each run is checked against the oracle on a few faces.
usage: run_configs.py [out.json]"""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import face_alignment_cvpr_2012_b200 as crf  # noqa: E402
from face_alignment_cvpr_2012_b200 import workloads as wl  # noqa: E402
from oracle import oracle as O  # noqa: E402


def timed(fn, reps=3):
    fn()
    t = []
    for _ in range(reps):
        t0 = time.perf_counter(); out = fn(); t.append(time.perf_counter() - t0)
    return out, float(np.median(t))


def main():
    mp = str(wl.staged_model_path())
    gm, om = crf.Model(packed=mp), O.Model(packed=mp)
    res = {}
    ctx = crf.Context(gm, 0)                                     # reference default strides 4 / 3
    # C1: the 20 LFW faces
    faces = wl.load_lfw()
    if faces:
        def c1():
            return [ctx.analyze_faces(f["img"], [f["box"]])[0] for f in faces]
        out, dt = timed(c1)
        ok = all(np.array_equal(out[i]["ffd"], om.analyze_face(faces[i]["img"], faces[i]["box"])["ffd"]) for i in (0, 7, 19))
        res["C1 eval_ffd on data/imgs (20 faces, one call per image)"] = {"faces_per_s": len(faces) / dt, "ms_per_face": 1e3 * dt / len(faces), "parity_spot_check": ok}
    # C3: 1080p frames, 16 faces each, one batched call
    frames, boxes, iob, _ = wl.make_frames(64, 1080, 1920, 16, seed=2013)
    out, dt = timed(lambda: ctx.analyze_batch(frames, boxes, iob))
    ok = all(np.array_equal(out[i]["ffd"], om.analyze_face(frames[iob[i]], boxes[i])["ffd"]) for i in (0, 100, len(boxes) - 1))
    res["C3 64 x 1080p frames, 16 faces/frame, boxes given (one crf_analyze_batch call, host frames)"] = {
        "faces": int(len(boxes)), "faces_per_s": len(boxes) / dt, "frames_per_s": 64 / dt, "parity_spot_check": ok}
    _, dt1 = timed(lambda: ctx.analyze_faces(frames[0], boxes[iob == 0]), reps=10)
    res["C3 one frame at a time (16 faces per crf_analyze_faces call)"] = {"ms_per_frame": 1e3 * dt1, "faces_per_s": int((iob == 0).sum()) / dt1}
    # C4: head pose only on 65536 crops (8192 distinct crops, tiled)
    base, _ = wl.make_crops(8192, seed=2014)
    crops = np.tile(base, (8, 1, 1, 1))
    out, dt = timed(lambda: ctx.analyze_crops(crops, headpose_only=True), reps=2)
    ok = all(out[i]["headpose"] == om.analyze_face(crops[i], (0, 0, 100, 100), headpose_only=True)["headpose"] for i in (0, 7, 65535))
    res["C4 head-pose forest only, 65536 crops, stride 4 (host crops)"] = {"faces_per_s": 65536 / dt, "parity_spot_check": ok}
    # C5: 64 mixed-resolution images
    imgs, _ = wl.make_mixed(64, seed=2015)
    def c5():
        return [ctx.analyze_faces(fr, bx) for fr, bx in imgs]
    out, dt = timed(c5, reps=2)
    n = sum(len(bx) for _, bx in imgs)
    fr, bx = imgs[9]
    ok = bool(np.array_equal(out[9][0]["ffd"], om.analyze_face(fr, bx[0])["ffd"]))
    res["C5 64 mixed-resolution images up to 4K (one call per image, host frames)"] = {"faces": n, "faces_per_s": n / dt, "images_per_s": 64 / dt, "parity_spot_check": ok}
    # C2 at the reference's default strides for context
    crops2, _ = wl.make_crops(4096, seed=2012)
    out, dt = timed(lambda: ctx.analyze_crops(crops2))
    res["C2 crops at the reference default strides 4/3 (4096 crops, host crops)"] = {"faces_per_s": 4096 / dt}
    print(json.dumps(res, indent=1))
    if len(sys.argv) > 1:
        Path(sys.argv[1]).write_text(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
