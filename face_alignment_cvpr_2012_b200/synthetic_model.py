"""Writer for random forests in the reference's on-disk format (Boost text archive v10).

Field orders follow the reference's serialize() methods (include/Tree.hpp:334-343,
include/Constants.hpp:44-59, include/TreeNode.hpp:148-164, include/ThresholdSplit.hpp:60-67,
include/ImageSample.hpp:83-90, include/opencv_serialization.hpp:65-79,
include/HeadPoseSample.hpp:154-161, include/MPSample.hpp:149-158); the token layout was checked
against the shipped data/trees_* files.  Used by the tests (loader edge cases, parity on forests
that are not the shipped ones) and as the model of last resort when staged/model.crfb200 is absent.
"""
from __future__ import annotations

from pathlib import Path

import numpy as np


def _fmt(v: float) -> str:
    return repr(float(np.float32(v))) if v != int(v) else str(int(v))


class _TreeWriter:
    def __init__(self, kind: str, max_depth: int, rng: np.random.Generator, channels: int, leaf_prob: float, max_rect: int, always_vote: bool = False):
        self.kind, self.max_depth, self.rng, self.channels, self.leaf_prob, self.max_rect = kind, max_depth, rng, channels, leaf_prob, max_rect
        self.always_vote = always_vote
        self.tok: list[str] = []
        self.oid = 0
        self.first = dict(node=True, split=True, leaf=True)

    def rect(self):
        w = int(self.rng.integers(1, self.max_rect + 1)); h = int(self.rng.integers(1, self.max_rect + 1))
        x = int(self.rng.integers(0, 31 - w)); y = int(self.rng.integers(0, 31 - h))  # x + w <= 30 like the shipped trees
        return [x, y, w, h]

    def node(self, depth: int) -> None:
        t = self.tok
        t.append("3")
        if self.first["node"]:
            t += ["1", "0"]; self.first["node"] = False
        t.append(str(self.oid)); self.oid += 1
        is_leaf = depth >= self.max_depth or (depth >= 2 and self.rng.random() < self.leaf_prob)
        t += [str(depth), "1" if is_leaf else "0", "0" if is_leaf else "1"]
        if is_leaf:
            if self.first["leaf"]:
                t += ["0", "0"]
            if self.kind == "hp":
                labels = self.rng.integers(0, 12, 5)
                n = int(labels.sum()) + int(self.rng.integers(0, 8))
                n = max(n, 1)
                fg = float(np.float32(labels.sum() / n))
                t += [str(n), _fmt(fg), "5", "0"] + [str(int(v)) for v in labels]
            else:
                n = int(self.rng.integers(3 if self.always_vote else 1, 40))
                t.append(str(n))
                if self.first["leaf"]:
                    t += ["0", "0"]  # class info of vector<Point>
                t += ["10", "0"]
                if self.first["leaf"]:
                    t += ["0", "0"]  # class info of Point
                t += [str(int(v)) for v in self.rng.integers(-60, 61, 20)]
                if self.always_vote:   # every leaf votes for every part (exercises the vote-budget overflow path)
                    t += ["10", "0"] + [_fmt(v) for v in self.rng.uniform(0.5, 20, 10)]
                    t += ["10", "0"] + [_fmt(v) for v in self.rng.uniform(0.5, 1, 10)]
                    t.append(_fmt(float(self.rng.choice([0.75, 1.0]))))
                else:
                    t += ["10", "0"] + [_fmt(v) for v in self.rng.uniform(0.5, 45, 10)]
                    t += ["10", "0"] + [_fmt(v) for v in self.rng.uniform(0, 1, 10) ** 2]
                    t.append(_fmt(float(self.rng.choice([0.0, 0.25, 0.5, 0.6, 0.75, 1.0, float(self.rng.uniform(0, 1))]))))
            self.first["leaf"] = False
        else:
            if self.first["split"]:
                t += ["0", "0", "0", "0"]  # ThresholdSplit, SimplePatchFeature
            t.append(str(int(self.rng.integers(0, self.channels))))
            if self.first["split"]:
                t += ["0", "0"]  # Rect
            self.first["split"] = False
            t += [str(v) for v in self.rect()] + [str(v) for v in self.rect()]
            t.append(repr(float(-self.rng.uniform(0.1, 2.5))))
            t.append(str(int(self.rng.integers(-60, 61))))
            self.node(depth + 1)
            self.node(depth + 1)


def tree_text(kind: str, max_depth: int, rng: np.random.Generator, channels: int = 38, leaf_prob: float = 0.15, max_rect: int = 22,
              ntrees: int = 15, face_size: int = 125, ratio: str = "0.25", finished: bool = True, always_vote: bool = False,
              features=(0, 1, 2)) -> str:
    w = _TreeWriter(kind, max_depth, rng, channels, leaf_prob, max_rect, always_vote)
    w.node(0)
    n_nodes = 2 ** (max_depth + 1) - 1
    path = "data/trees_headpose" if kind == "hp" else "data/trees_ffd"
    hdr = ["22", "serialization::archive", "10", "0", "0", str(n_nodes), str(n_nodes if finished else n_nodes - 1), "0", "0",
           str(max_depth), "20", "2000", str(ntrees), "400", "200", str(face_size), ratio, str(len(path)), path, "9", "index.txt", str(len(features)), "0"] + [str(int(f)) for f in features] + [
           "8", "tree.txt"]
    return " ".join(hdr + w.tok) + "\n"


def write_forest(dir_: Path, kind: str, ntrees: int, max_depth: int, seed: int, **kw) -> None:
    dir_ = Path(dir_)
    dir_.mkdir(parents=True, exist_ok=True)
    for i in range(ntrees):
        rng = np.random.default_rng([seed, i])
        (dir_ / f"tree_{i:03d}.txt").write_text(tree_text(kind, max_depth, rng, ntrees=ntrees, **kw))


def write_model(root: Path, seed: int = 7, hp_trees: int = 15, hp_depth: int = 10, ffd_trees: int = 20, ffd_depth: int = 11, **kw):
    """Writes root/trees_headpose and root/trees_ffd/forest_0..4; returns (hp_dir, ffd_dir)."""
    root = Path(root)
    write_forest(root / "trees_headpose", "hp", hp_trees, hp_depth, seed, **kw)
    for f in range(5):
        write_forest(root / "trees_ffd" / f"forest_{f}", "mp", ffd_trees, ffd_depth, seed * 100 + f + 1, **kw)
    return str(root / "trees_headpose"), str(root / "trees_ffd")
