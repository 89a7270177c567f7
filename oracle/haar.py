"""CPU restatement of the face-box source of FaceForest::detectFace (src/FaceForest.cpp:136-159):
cv::CascadeClassifier::detectMultiScale on data/haarcascade_frontalface_alt.xml (stump-based HAAR cascade, new XML format)
followed by the reference's box enlargement.  TEST INFRASTRUCTURE ONLY (oracle for the GPU cascade evaluator).

The algorithm lives in OpenCV (objdetect/cascadedetect.cpp), which is absent from /root/reference; restated from its published
behaviour for the call the reference makes (scaleFactor 1.3, minNeighbors 1, flags 0, minSize 30x30):
  scales   factor = 1, 1.3, 1.3^2, ...; window = round(20 * factor); skipped while window < minSize, stops when the window exceeds
           the image or the scaled image is smaller than the 20x20 window
  pyramid  gray image resized to round(W / factor) x round(H / factor) (bilinear), integral of values and of squares per level
  window   variance normalisation over the inner 18x18 rect: nf = sqrt(area * sqsum - sum^2), windows with nf <= 0 are skipped
           (OpenCV additionally rejects area / nf >= 0.1)
  stump    feature = sum_k weight_k * rectsum_k ; leaf = value * (1 / nf) < threshold ? left : right ; stage passes if the sum of its
           leaves >= stageThreshold; ystep = 2 for factor < 2 else 1, one extra step after a first-stage reject
  grouping cv::groupRectangles(minNeighbors, eps 0.2): union of similar rects, classes with <= minNeighbors members dropped, then
           small rects inside better-supported larger ones dropped
Pinned statistically to cv2 4.13's detectMultiScale on the 20 shipped LFW images (tests/test_haar.py: same number of boxes,
IoU >= 0.9): cv2 builds its pyramid with INTER_LINEAR_EXACT, whose fixed-point rounding differs from INTER_LINEAR in the last
bit of a few pixels, so single marginal windows may differ while the grouped boxes agree.
"""
from __future__ import annotations

import xml.etree.ElementTree as ET
from pathlib import Path

import numpy as np


class Cascade:
    def __init__(self, path: str | Path):
        root = ET.parse(str(path)).getroot().find("cascade")
        self.w, self.h = int(root.find("width").text), int(root.find("height").text)
        feats = []
        for f in root.find("features"):
            rects = [[float(v) for v in r.text.split()] for r in f.find("rects")]
            assert f.find("tilted") is None or int(f.find("tilted").text) == 0
            while len(rects) < 3:
                rects.append([0, 0, 0, 0, 0.0])
            feats.append(rects)
        self.features = np.array(feats, np.float64)          # [nfeat][3][x, y, w, h, weight]
        self.stages = []
        for st in root.find("stages"):
            thr = float(st.find("stageThreshold").text)
            weak = []
            for wc in st.find("weakClassifiers"):
                node = wc.find("internalNodes").text.split()
                leaves = [float(v) for v in wc.find("leafValues").text.split()]
                assert int(node[0]) == 0 and int(node[1]) == -1 and len(leaves) == 2     # stumps
                weak.append((int(node[2]), float(node[3]), leaves[0], leaves[1]))
            self.stages.append((np.float32(thr), weak))

    def flat(self):
        """Arrays for the device image: per stage (first weak, count, threshold), per weak (feature, threshold, left, right),
        per feature 3 x (x, y, w, h, weight)."""
        st, wk = [], []
        for thr, weak in self.stages:
            st.append((len(wk), len(weak), float(thr)))
            wk += weak
        return np.array(st, np.float64), np.array(wk, np.float64), self.features


def _resize(gray, dh, dw):
    from . import oracle as O
    return O.resize(gray, dh, dw)


def detect_candidates(c: Cascade, gray: np.ndarray, scale_factor: float = 1.3, min_size: int = 30):
    H, W = gray.shape
    out = []
    factor = 1.0
    while True:
        win = (int(round(c.w * factor)), int(round(c.h * factor)))
        sw, sh = int(round(W / factor)), int(round(H / factor))
        if win[0] > W or win[1] > H or sw < c.w or sh < c.h:
            break
        if win[0] >= min_size and win[1] >= min_size:
            img = _resize(gray, sh, sw).astype(np.int64) if factor != 1.0 else gray.astype(np.int64)
            S = np.zeros((sh + 1, sw + 1), np.int64); S[1:, 1:] = img.cumsum(0).cumsum(1)
            Q = np.zeros((sh + 1, sw + 1), np.int64); Q[1:, 1:] = (img * img).cumsum(0).cumsum(1)
            ystep = 1 if factor > 2.0 else 2
            ny, nx = sh - c.h + 1, sw - c.w + 1
            ys, xs = np.arange(0, ny, ystep), np.arange(0, nx, ystep)
            # setWindow rejects x + w >= szi.width (= sw + 1), i.e. keeps x <= sw - w
            yy, xx = np.meshgrid(ys, xs, indexing="ij")

            def rs(A, x, y, w, h):
                return A[yy + y + h, xx + x + w] - A[yy + y, xx + x + w] - A[yy + y + h, xx + x] + A[yy + y, xx + x]

            area = float((c.w - 2) * (c.h - 2))
            vs = rs(S, 1, 1, c.w - 2, c.h - 2).astype(np.float64)
            vq = (rs(Q, 1, 1, c.w - 2, c.h - 2) & 0xffffffff).astype(np.float64)     # OpenCV keeps the square sums in 32 bits
            nf = area * vq - vs * vs
            ok = nf > 0
            inv = np.where(ok, 1.0 / np.sqrt(np.where(ok, nf, 1.0)), 1.0).astype(np.float32)
            ok &= (np.float32(area) * inv) < np.float32(0.1)
            alive = ok.copy()
            first_fail = np.zeros_like(alive)
            for si, (thr, weak) in enumerate(c.stages):
                if not alive.any():
                    break
                ssum = np.zeros(alive.shape, np.float32)
                for fi, t, left, right in weak:
                    f = c.features[fi]
                    v = np.zeros(alive.shape, np.float32)
                    for k in range(3):
                        x, y, w, h, wt = f[k]
                        if wt != 0:
                            v = v + np.float32(wt) * rs(S, int(x), int(y), int(w), int(h)).astype(np.float32)
                    ssum = ssum + np.where(v * inv < np.float32(t), np.float32(left), np.float32(right))
                passed = ssum >= thr
                if si == 0:
                    first_fail = alive & ~passed
                alive &= passed
            # the scan skips one extra step after a first-stage reject: x += ystep when result == 0
            for iy in range(len(ys)):
                ix = 0
                while ix < len(xs):
                    if alive[iy, ix]:
                        out.append((int(round(xs[ix] * factor)), int(round(ys[iy] * factor)), win[0], win[1]))
                    if first_fail[iy, ix]:   # runAt() == 0 (rejected by the first stage): x += ystep once more; an invalid window returns -1
                        ix += 1
                    ix += 1
        factor *= scale_factor
    return out


def group_rectangles(rects, group_threshold: int = 1, eps: float = 0.2):
    """cv::groupRectangles (objdetect/cascadedetect.cpp)."""
    n = len(rects)
    if group_threshold <= 0 or n == 0:
        return list(rects)
    parent = list(range(n))

    def find(i):
        while parent[i] != i:
            parent[i] = parent[parent[i]]
            i = parent[i]
        return i

    def similar(a, b):
        d = eps * (min(a[2], b[2]) + min(a[3], b[3])) * 0.5
        return abs(a[0] - b[0]) <= d and abs(a[1] - b[1]) <= d and abs(a[0] + a[2] - b[0] - b[2]) <= d and abs(a[1] + a[3] - b[1] - b[3]) <= d

    for i in range(n):
        for j in range(i + 1, n):
            if similar(rects[i], rects[j]):
                a, b = find(i), find(j)
                if a != b:
                    parent[b] = a
    roots = {}
    labels = []
    for i in range(n):
        r = find(i)
        labels.append(roots.setdefault(r, len(roots)))
    nc = len(roots)
    acc = np.zeros((nc, 4), np.int64); cnt = np.zeros(nc, np.int64)
    for i, l in enumerate(labels):
        acc[l] += rects[i]; cnt[l] += 1
    rr = []
    for l in range(nc):
        s = np.float32(1.0) / np.float32(cnt[l])
        rr.append(tuple(int(np.rint(np.float32(v) * s)) for v in acc[l]))
    keep = []
    for i in range(nc):
        if cnt[i] <= group_threshold:
            continue
        r1, n1 = rr[i], cnt[i]
        for j in range(nc):
            n2 = cnt[j]
            if j == i or n2 <= group_threshold:
                continue
            r2 = rr[j]
            dx, dy = int(np.rint(r2[2] * eps)), int(np.rint(r2[3] * eps))
            if r1[0] >= r2[0] - dx and r1[1] >= r2[1] - dy and r1[0] + r1[2] <= r2[0] + r2[2] + dx and r1[1] + r1[3] <= r2[1] + r2[3] + dy and (n2 > max(3, n1) or n1 < 3):
                break
        else:
            keep.append(r1)
    return keep


def detect_multi_scale(c: Cascade, bgr_or_gray: np.ndarray, scale_factor: float = 1.3, min_neighbors: int = 1, min_size: int = 30):
    from . import oracle as O
    gray = O.bgr2gray(bgr_or_gray) if bgr_or_gray.ndim == 3 else bgr_or_gray
    return group_rectangles(detect_candidates(c, gray, scale_factor, min_size), min_neighbors)


def enlarge(boxes, rows: int, cols: int):
    """src/FaceForest.cpp:147-157: 5 % of the width left and right, 15 % of the WIDTH added to the height twice, clipped."""
    out = []
    for (x, y, w, h) in boxes:
        ox, oy = int(w * 0.05), int(w * 0.15)
        x0, y0, x1, y1 = max(x - ox, 0), max(y, 0), min(x - ox + w + 2 * ox, cols), min(y + h + 2 * oy, rows)
        out.append((x0, y0, x1 - x0, y1 - y0) if x1 > x0 and y1 > y0 else (0, 0, 0, 0))
    return out
