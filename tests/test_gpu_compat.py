"""GPU: the reference-shaped C++ interface (include/crf_b200_compat.hpp) — every class and free function of SURVEY §8(b),
driven by tests/cpp/compat_full.cc in the order FaceForest::analyzeFace uses them, against the oracle.  Built twice: with the
cvlite stand-ins and with CRF_B200_WITH_OPENCV against a stub <opencv2/core/core.hpp> (so that branch cannot rot)."""
import subprocess
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


def _leaf(om, which, tree, oid):
    d = om.leaf_dump(which, tree)
    row = d[d[:, 0] == oid]
    assert len(row) == 1
    return row[0]


@pytest.mark.parametrize("opencv", [False, True])
def test_reference_interface_in_cpp(crf, O, synth_dirs, synth_models, tmp_path, opencv):
    if crf.lib().crf_device_count() < 1:
        pytest.fail("no CUDA device: the gpu-marked tests must run on the B200 box (there is no CPU fallback)")
    from face_alignment_cvpr_2012_b200 import capi
    exe = tmp_path / "compat_full"
    flags = ["-DCRF_B200_WITH_OPENCV", "-I", str(ROOT / "tests" / "cpp" / "opencv_stub")] if opencv else []
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", *flags, "-I", str(ROOT / "include"), str(ROOT / "tests" / "cpp" / "compat_full.cc"), "-o", str(exe),
                    "-L", str(capi.LIB_PATH.parent), "-lcrf_b200", f"-Wl,-rpath,{capi.LIB_PATH.parent}"], check=True)
    hp, ffd = synth_dirs
    from face_alignment_cvpr_2012_b200 import workloads as wl
    extra = []
    xml = wl.STAGED / "haarcascade_frontalface_alt.xml"
    lfw = wl.load_lfw()
    if xml.exists() and lfw:
        import cv2
        cv2.imwrite(str(tmp_path / "face.ppm"), lfw[0]["img"])
        extra = [str(xml), str(tmp_path / "face.ppm")]
    r = subprocess.run([str(exe), hp, ffd, *extra], capture_output=True, text=True)
    assert r.returncode == 0 and "compat_full ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    out = {}
    for line in r.stdout.split("\n"):
        t = line.split()
        if t and "." in t[0] or (t and t[0] in ("composed", "rescaled", "areaUnderCurve", "votes", "estimateFacialFeatures", "estimateHeadPose")):
            out[t[0]] = t[1:]
    _, om = synth_models
    px = ((np.arange(120 * 160 * 3, dtype=np.uint64) * 2654435761 % (1 << 32)) >> 24).astype(np.uint8).reshape(120, 160, 3)
    box = (10, 5, 100, 100)
    want = om.analyze_face(px, box)
    # FaceForest::analyzeFace
    assert np.float32(out["analyzeFace.headpose"][0]) == want["headpose"]
    assert [int(v) for v in out["analyzeFace.ffd"]] == want["ffd"].ravel().tolist()
    counts, dom, fi, ti, _ = om.compose(want["headpose"], want["variance"])
    composed = [f"{a}:{b}" for a, b in zip(fi.tolist(), ti.tolist())]
    assert out["analyzeFace.composed"] == composed
    # ImageSample
    x, y, w, h = box
    sw, sh, _ = O.scaled_size(w, h)
    g = O.resize(O.bgr2gray(px)[y:y + h, x:x + w], sh, sw)
    planes, integ = O.channels(g)
    assert out["sample.channels"] == ["38"]
    corners = [l.split() for l in r.stdout.split("\n") if l.startswith("sample.corner")]
    assert [(int(c[1]), float(c[2])) for c in corners] == [(c, float(integ[c][sh][sw])) for c in (0, 5, 20, 37)]
    I = integ.astype(np.int64)
    def mean(c, px_, py_, rx, ry, rw, rh):
        return int(I[c, py_ + ry + rh, px_ + rx + rw] - I[c, py_ + ry, px_ + rx + rw] - I[c, py_ + ry + rh, px_ + rx] + I[c, py_ + ry, px_ + rx]) // (rw * rh)
    assert int(out["sample.evalTest"][0]) == mean(9, 20, 30, 3, 4, 10, 7) - mean(9, 20, 30, 12, 15, 5, 13)
    assert int(out["plain.type8u"][0]) == 1 and int(out["plain.evalTest"][0]) == int(out["sample.evalTest"][0])
    # Forest<HeadPoseSample>::load / evaluateMT, Tree::evaluateMT, TreeNode
    s = O.Sample(planes=planes)
    ny1 = sh - 31
    ids1, _, _, _ = om.eval_hp(s, 1)
    assert out["hp_forest.trees"] == ["15", "patch", "31"]
    got = out["hp_forest.evaluateMT"]
    for t in range(15):
        L = _leaf(om, -1, t, ids1[12 * ny1 + 20][t])
        n, fg, lab = got[t].split(":")
        assert int(n) == int(L[1]) and np.float32(fg) == np.float32(f"{L[2]:.6g}") and int(lab) == int(L[5])
    assert out["tree.evaluateMT.root"] == ["1"] and out["tree.evaluateMT.inner"] == ["1"]
    root = om.tree_dump(-1, 3)[0]
    assert [int(v) for v in out["tree.root.split"]] == [int(root[2]), int(root[3]), int(root[5]), int(root[8]), int(root[10]), int(root[11])]
    # Tree<MPSample>::load + evaluateMT
    e1 = om.eval_ffd(s, [1], [4], 1)
    L = _leaf(om, 1, 4, e1["leaf_ids"][40 * ny1 + 50][0])
    m = out["mp_tree.evaluateMT"]
    assert int(m[0]) == int(L[1]) and np.float32(m[1]) == np.float32(f"{L[2]:.6g}") and (int(m[2]), int(m[3])) == (int(L[3 + 6]), int(L[4 + 6]))
    assert np.float32(m[4]) == np.float32(f"{L[33 + 7]:.6g}") and out["mp_tree.load.missing"] == ["0"]
    # estimateHeadPose / getHeadPoseVotesMT / areaUnderCurve / composition
    _, hp4, var4, _ = om.eval_hp(s, 4)
    assert np.float32(out["estimateHeadPose"][0]) == hp4 and np.float32(out["estimateHeadPose"][1]) == var4
    assert hp4 == want["headpose"] and var4 == want["variance"]
    _, hp2, var2, _ = om.eval_hp(s, 2)
    assert np.float32(out["getHeadPoseVotesMT.step2"][0]) == hp2 and np.float32(out["getHeadPoseVotesMT.step2"][1]) == var2
    T = np.array([-2.5, -0.35, -0.20, 0.20, 0.35, 2.5], np.float32)
    areas = [O.area_under_curve(float(T[j]), float(T[j + 1]), float(hp4), float(np.sqrt(np.float64(var4)))) for j in range(5)]
    assert [np.float32(v) for v in out["areaUnderCurve"]] == areas
    assert out["composed"] == composed
    # estimateFacialFeatures / getFacialFeaturesVotesMT / MeanShift::shift / Forest<MPSample>::evaluateMT
    e = om.eval_ffd(s, fi, ti, 3, vote_cap=20000)
    assert [int(v) for v in out["estimateFacialFeatures"]] == e["rounded"].ravel().tolist()
    assert [int(v) for v in out["rescaled"]] == want["ffd"].ravel().tolist()
    for p in range(10):
        n = int(e["n_votes"][p]); hsh = 0
        for k in range(n):
            vx, vy, vw = e["votes"][p][k]
            hsh = (hsh * 31 + (int(vx) + 1000) * 7 + (int(vy) + 1000) + int(np.float32(vw) * np.float32(1024))) % 1000000007
        assert out["votes"][p] == f"{n}:{hsh}", p
    assert [int(v) for v in out["MeanShift.shift"]] == e["rounded"].ravel().tolist()
    ef = om.eval_ffd(s, fi, ti, 1)
    ids = ef["leaf_ids"][33 * ny1 + 41]
    assert [int(v) for v in out["mp_forest.evaluateMT"]] == [int(_leaf(om, int(fi[k]), int(ti[k]), ids[k])[1]) for k in range(len(fi))]
    assert out["multi.same"] == ["1"]
    if extra:   # analyzeImage(img, faces): GPU cascade -> enlarged box -> analyzeFace, against the oracle on the same box
        from oracle import haar as H
        d = out["detect.faces"]
        assert d[0] == "1" and out["detect.badcascade"] == ["0"]
        box = tuple(int(v) for v in d[1:5])
        want_box = H.enlarge(H.detect_multi_scale(H.Cascade(str(xml)), lfw[0]["img"]), *lfw[0]["img"].shape[:2])[0]
        assert box == want_box
        w2 = om.analyze_face(lfw[0]["img"], box)
        assert np.float32(d[5]) == w2["headpose"] and (int(d[6]), int(d[7])) == tuple(w2["ffd"][0])
    s.close()
