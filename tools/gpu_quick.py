#!/usr/bin/env python
"""Quick stage-by-stage parity + timing probe on a GPU box (development aid; tests/ holds the real checks)."""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import face_alignment_cvpr_2012_b200 as crf  # noqa: E402
from face_alignment_cvpr_2012_b200 import workloads as wl  # noqa: E402
from oracle import oracle as O  # noqa: E402


def main():
    n_time = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    mp = str(wl.staged_model_path())
    gm = crf.Model(packed=mp)
    om = O.Model(packed=mp)
    faces = wl.load_lfw()
    print("model", gm.info, "lfw faces", len(faces))
    ctx = crf.Context(gm, 0)
    bad = 0
    for f in faces[:4]:
        sc = ctx.stage_gray_resize(f["img"], f["box"])
        x, y, w, h = f["box"]
        g = O.bgr2gray(f["img"])[y:y + h, x:x + w]
        sw, sh, _ = O.scaled_size(w, h)
        ref = O.resize(g, sh, sw)
        d = int((sc.astype(int) != ref.astype(int)).sum())
        print(f["name"], "resize mismatches", d, sc.shape)
        bad += d
        planes, integ = ctx.stage_channels(ref)
        op, oi = O.channels(ref)
        dp = (planes != op).reshape(38, -1).sum(1)
        di = int((integ != oi.astype(np.uint32)).sum())
        print("  plane mismatches per plane:", dp.tolist() if dp.any() else 0, "integral mismatches", di)
        bad += int(dp.sum()) + di
        s = O.Sample(planes=op)
        ids_o, hp_o, var_o, vis = om.eval_hp(s, 4)
        ids_g = ctx.stage_eval_forest(op, 4)
        print("  hp leaf mismatches", int((ids_o != ids_g).sum()), "of", ids_o.size)
        bad += int((ids_o != ids_g).sum())
        r = ctx.stage_headpose(op, 4)
        counts, dom, fi, ti, fl = om.compose(hp_o, var_o)
        print("  headpose", r["headpose"], hp_o, "var", r["variance"], var_o, "counts", r["tree_counts"].tolist(), counts.tolist())
        bad += int(r["headpose"] != hp_o) + int(r["variance"] != var_o) + int((r["tree_counts"] != counts).any())
        e = om.eval_ffd(s, fi, ti, 3, vote_cap=4096)
        ids_g = ctx.stage_eval_forest(op, 3, fi, ti)
        print("  ffd leaf mismatches", int((e["leaf_ids"] != ids_g).sum()), "of", ids_g.size)
        bad += int((e["leaf_ids"] != ids_g).sum())
        v = ctx.stage_votes_meanshift(op, 3, fi, ti, vote_cap=4096)
        print("  n_votes eq", bool((v["n_votes"] == e["n_votes"]).all()), "votes eq", bool((v["votes"] == e["votes"]).all()),
              "max |mean diff|", float(np.abs(v["mean"] - e["mean"]).max()), "iters", v["iters"].tolist(), e["iters"].tolist())
        bad += int((v["n_votes"] != e["n_votes"]).any()) + int((v["votes"] != e["votes"]).any())
        s.close()
    # whole path on the 20 faces
    for f in faces:
        g = ctx.analyze_faces(f["img"], [f["box"]])[0]
        o = om.analyze_face(f["img"], f["box"])
        ok = g["headpose"] == o["headpose"] and (g["tree_counts"] == o["tree_counts"]).all() and (g["n_votes"] == o["n_votes"]).all()
        dm = float(np.abs(g["ffd_f"] - o["ffd_f"]).max())
        print(f["name"], "ok" if ok else "MISMATCH", "max|ffd_f diff|", dm, "ffd eq", bool((g["ffd"] == o["ffd"]).all()))
        bad += int(not ok)
    print("TOTAL BAD", bad)
    # timing, stride 1
    crops, tag = wl.make_crops(n_time)
    opt = crf._options(None, hp_stride=1, ffd_stride=1)
    ctx1 = crf.Context(gm, 0, opt)
    ctx1.analyze_crops(crops[:64])
    ctx1.set_profiling(True, True)
    ctx1.reset_counters()
    t = time.time()
    out = ctx1.analyze_crops(crops)
    dt = time.time() - t
    print(f"stride-1: {n_time} crops in {dt * 1e3:.1f} ms -> {n_time / dt:.0f} faces/s")
    print("stage ms", ctx1.stage_ms())
    print("counters", ctx1.counters())
    o = om.analyze_face(crops[0], (0, 0, 100, 100), 1, 1)
    print("stride-1 face0 parity: hp", out[0]["headpose"], o["headpose"], "max|ffd_f diff|", float(np.abs(out[0]["ffd_f"] - o["ffd_f"]).max()),
          "votes eq", bool((out[0]["n_votes"] == o["n_votes"]).all()))
    ctx1.set_profiling(False, False)
    for _ in range(2):
        t = time.time(); ctx1.analyze_crops(crops); dt = time.time() - t
        print(f"stride-1 (no profiling): {n_time / dt:.0f} faces/s")
    # default strides
    ctx.analyze_crops(crops[:64])
    ctx.set_profiling(True, True); ctx.reset_counters()
    t = time.time(); ctx.analyze_crops(crops); dt = time.time() - t
    print(f"default strides: {n_time / dt:.0f} faces/s", ctx.stage_ms()[0])


if __name__ == "__main__":
    main()
