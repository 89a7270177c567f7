#!/usr/bin/env python
"""Large randomized parity campaign: the GPU path vs the CPU oracle on thousands of faces (evidence for DESIGN.md / profiles/).
usage: parity_campaign.py [n_default=3000] [n_stride1=64] [out.json]"""
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


def compare(got, want):
    exact = lambda k: int((got[k] != want[k]).reshape(len(got), -1).any(axis=1).sum())  # noqa: E731
    hp_bits = int((got["headpose"].view(np.uint32) != want["headpose"].view(np.uint32)).sum())
    var_bits = int((got["variance"].view(np.uint32) != want["variance"].view(np.uint32)).sum())
    d = np.abs(got["ffd_f"] - want["ffd_f"])
    return {"faces": int(len(got)), "headpose_bit_mismatch": hp_bits, "variance_bit_mismatch": var_bits,
            "tree_counts_mismatch": exact("tree_counts"), "n_votes_mismatch": exact("n_votes"), "ms_iters_mismatch": exact("ms_iters"),
            "max_abs_ffd_f_diff_px": float(np.nanmax(d)), "faces_with_any_ffd_f_diff": int((d.reshape(len(got), -1) > 0).any(axis=1).sum()),
            "integer_landmarks_differing": int((got["ffd"] != want["ffd"]).sum()), "integer_landmarks_total": int(got["ffd"].size)}


def main():
    n_def = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
    n_s1 = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    mp = str(wl.staged_model_path())
    gm, om = crf.Model(packed=mp), O.Model(packed=mp)
    cores = O.hardware_concurrency()
    res = {}
    crops, tag = wl.make_crops(n_def, seed=777)
    t = time.time()
    got = crf.Context(gm, 0).analyze_crops(crops)
    want = np.array([om.analyze_face(c, (0, 0, 100, 100), threads=cores) for c in crops])
    res["default strides 4/3, 100x100 crops (incl. every 8th pure noise)"] = compare(got, want)
    # ragged boxes in frames
    frames, boxes, iob, _ = wl.make_frames(24, 720, 1280, 12, seed=99, wmin=64, wmax=400)
    got = crf.Context(gm, 0).analyze_batch(frames, boxes, iob)
    want = np.array([om.analyze_face(frames[i], b, threads=cores) for b, i in zip(boxes, iob)])
    res["default strides, ragged boxes in 720p frames"] = compare(got, want)
    crops1 = crops[:n_s1]
    got = crf.Context(gm, 0, crf._options(None, hp_stride=1, ffd_stride=1)).analyze_crops(crops1)
    want = np.array([om.analyze_face(c, (0, 0, 100, 100), 1, 1, threads=cores) for c in crops1])
    res["stride 1/1 (BASELINE config 2 options)"] = compare(got, want)
    res["seconds"] = time.time() - t
    res["data"] = tag
    print(json.dumps(res, indent=1))
    if len(sys.argv) > 3:
        Path(sys.argv[3]).write_text(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
