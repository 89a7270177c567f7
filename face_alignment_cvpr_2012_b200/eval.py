"""eval_ffd / eval_headpose drivers on top of the GPU library.

Mirrors the reference's evaluation mains: src/eval_ffd.cpp:57-179 (evalForest, getInterOccularDist, 90/10 split per pose
class, output/errors.txt) and src/eval_headpose.cpp:57-139, with loadAnnotations / loadImage of
src/face_utils.cpp:18-28,142-181 and loadConfigFile (:50-140).  Image decoding is cv2.imread, as in the reference.

  python -m face_alignment_cvpr_2012_b200.eval ffd      --config-ffd data/config_ffd.txt --config-headpose data/config_headpose.txt
  python -m face_alignment_cvpr_2012_b200.eval headpose ...
Options: --all (evaluate every annotation instead of the last 10 % of each pose class), --packed MODEL (pre-packed forests),
--annotations FILE / --image-dir DIR (override the paths of the config file), --out output/errors.txt
"""
from __future__ import annotations

import argparse
import math
import sys
from dataclasses import dataclass, field
from pathlib import Path

import numpy as np

NUM_HEADPOSE_CLASSES = 5          # include/Constants.hpp:66
TRAIN_IMAGES_PERCENTAGE = 0.9     # include/Constants.hpp:68


@dataclass
class FaceAnnotation:             # include/face_utils.hpp:44-52
    url: str = ""
    bbox: tuple = (0, 0, 0, 0)
    pose: int = 0
    parts: np.ndarray = field(default_factory=lambda: np.zeros((0, 2), np.int32))


def loadAnnotations(path: str) -> list | None:
    """src/face_utils.cpp:142-181.  Returns None when the file does not exist (the reference returns false)."""
    p = Path(path)
    if not p.exists():
        return None
    print(f"Open annotations file: {path}")
    out = []
    for line in p.read_text().split("\n"):
        strs = line.split(" ")
        if not strs or strs[0] == "#" or len(strs) < 7:
            continue
        n = int(strs[6])
        parts = np.array([[int(strs[7 + 2 * i]), int(strs[8 + 2 * i])] for i in range(n)], np.int32)
        out.append(FaceAnnotation(strs[0], (int(strs[1]), int(strs[2]), int(strs[3]), int(strs[4])), int(strs[5]), parts))
    return out


def loadImage(path: str, name: str):
    """src/face_utils.cpp:18-28: the image lives next to the annotation file."""
    import cv2
    pos = path.rfind("/") + 1
    return cv2.imread(path[:pos] + name, cv2.IMREAD_COLOR)


def getInterOccularDist(ann: FaceAnnotation) -> float:
    """src/eval_ffd.cpp:33-46."""
    p = ann.parts.astype(np.float32)
    cl = (p[0] + p[1]) / np.float32(2.0)
    cr = (p[6] + p[7]) / np.float32(2.0)
    return float(np.float32(math.sqrt(float(cl[0] - cr[0]) ** 2 + float(cl[1] - cr[1]) ** 2)))


def split_test(annotations: list, everything: bool = False) -> list:
    """src/eval_ffd.cpp:155-169: group by pose class, keep the last 10 % of each class (class order, file order inside)."""
    by_pose = [[] for _ in range(NUM_HEADPOSE_CLASSES)]
    for a in annotations:
        by_pose[a.pose + 2].append(a)
    out = []
    for cls in by_pose:
        n_train = 0 if everything else int(len(cls) * TRAIN_IMAGES_PERCENTAGE)
        out += cls[n_train:]
    return out


def evalForest_ffd(ff, annotations: list, image_path: str, out_path: str | None = "output/errors.txt") -> np.ndarray:
    """src/eval_ffd.cpp:57-122.  Returns the error matrix [n, 10] it also writes to out_path."""
    errors = []
    for a in annotations:
        img = loadImage(image_path, a.url)
        if img is None:
            print(f"Could not load: {a.url}", file=sys.stderr)
            continue
        face = ff.analyzeFace(img, a.bbox)
        iod = getInterOccularDist(a)
        d = a.parts.astype(np.float64) - face.ffd_cordinates.astype(np.float64)
        errors.append([float(np.float32(math.sqrt(dx * dx + dy * dy)) / np.float32(iod)) for dx, dy in d])
    err = np.array(errors, np.float32).reshape(-1, 10)
    if out_path:
        Path(out_path).parent.mkdir(parents=True, exist_ok=True)
        with open(out_path, "w") as f:
            for row in err:
                f.write("".join(f"{v:g} " for v in row) + "\n")   # ofs << errors[i][j] << " " (:116-121)
    return err


def evalForest_headpose(ff, annotations: list, image_path: str) -> list:
    """src/eval_headpose.cpp:57-90: prints Real / Predict per image."""
    out = []
    for a in annotations:
        img = loadImage(image_path, a.url)
        if img is None:
            print(f"Could not load: {a.url}", file=sys.stderr)
            continue
        face = ff.analyzeFace(img, a.bbox)
        print(f"Real:{a.pose} Predict:{face.headpose:g}")
        out.append((a.pose, face.headpose))
    return out


def main(argv=None) -> int:
    from . import FaceForest, FaceForestOptions, loadConfigFile
    ap = argparse.ArgumentParser(prog="face_alignment_cvpr_2012_b200.eval")
    ap.add_argument("what", choices=["ffd", "headpose"])
    ap.add_argument("--config-ffd", default="data/config_ffd.txt")
    ap.add_argument("--config-headpose", default="data/config_headpose.txt")
    ap.add_argument("--packed", default="")
    ap.add_argument("--annotations", default="")
    ap.add_argument("--all", action="store_true")
    ap.add_argument("--out", default="output/errors.txt")
    args = ap.parse_args(argv)
    try:
        hp_param, mp_param = loadConfigFile(args.config_headpose), loadConfigFile(args.config_ffd)
    except FileNotFoundError as e:
        print(f"(!) {e}", file=sys.stderr)
        return 1
    opt = FaceForestOptions(hp_forest_param=hp_param, mp_forest_param=mp_param, packed_model=args.packed)
    ann_path = args.annotations or (mp_param.image_path if args.what == "ffd" else hp_param.image_path)
    annotations = loadAnnotations(ann_path)
    if annotations is None:
        print(f"(!) annotations file not found: {ann_path}", file=sys.stderr)
        return 1
    ff = FaceForest(opt)
    if not ff.is_inizialized:
        return 1
    test = split_test(annotations, args.all)
    if args.what == "ffd":
        err = evalForest_ffd(ff, test, ann_path, args.out)
        print(f"{len(err)} faces, mean error / inter-ocular distance per part: " + " ".join(f"{v:.4f}" for v in err.mean(axis=0)) + f"  (all parts {err.mean():.4f})")
    else:
        evalForest_headpose(ff, test, ann_path)
    return 0


if __name__ == "__main__":
    sys.exit(main())
