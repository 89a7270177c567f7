"""GPU: configurable feature lists, the per-sample level of the interface (evaluateMT on explicit patches,
ImageSample::evalTest) and contexts on partial models — each against the oracle."""
from pathlib import Path

import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu(crf):
    if crf.lib().crf_device_count() < 1:
        pytest.fail("no CUDA device: the gpu-marked tests must run on the B200 box (there is no CPU fallback)")
    return True


def _img(rng, H, W):
    import cv2
    return cv2.GaussianBlur(rng.integers(0, 256, (H, W), dtype=np.uint8), (0, 0), 1.5)


@pytest.mark.parametrize("features", [(0, 1, 2, 3, 4, 5), (3, 0), (5, 4, 2), (1,), (0, 2, 3)])
def test_feature_channel_lists(O, crf, gpu, features):
    """FeatureChannelFactory::extractChannel over a configured list: planes appended in sorted-id order."""
    ctx = crf.Context(None, 0)
    rng = np.random.default_rng(sum(features))
    for H, W in [(125, 125), (141, 124)]:
        img = _img(rng, H, W)
        planes, integ = ctx.stage_feature_channels(img, features)
        op, oi = O.channels(img, features_mask=sum(1 << f for f in features))
        assert planes.shape == op.shape and np.array_equal(planes, op)
        assert np.array_equal(integ, oi.astype(np.uint32))
    with pytest.raises(crf.CrfError):
        ctx.stage_feature_channels(img, [0, 0])


def test_forests_on_42_planes_end_to_end(O, crf, gpu, tmp_path):
    """A forest trained on all six feature kinds (42 planes; its splits read channels up to 41): whole path and stride-1
    (the shared-memory window holds at most 38 planes, so dense grids fall back to the global-gather kernels)."""
    from face_alignment_cvpr_2012_b200 import synthetic_model as sm, workloads as wl
    hp, ffd = sm.write_model(tmp_path / "all", seed=21, channels=42, features=(0, 1, 2, 3, 4, 5))
    gm, om = crf.Model(hp, ffd, 15, 20), O.Model(hp, ffd, 15, 20)
    used = max(int(gm.tree_dump(w, t)[:, 2].max()) for w in (-1, 0, 3) for t in range(5))
    assert used >= 38
    crops, _ = wl.make_crops(12, seed=5)
    for hs, fs, sel in [(4, 3, range(12)), (1, 1, (0, 7))]:
        ctx = crf.Context(gm, 0, crf._options(None, hp_stride=hs, ffd_stride=fs))
        got = ctx.analyze_crops(crops)
        for i in sel:
            want = om.analyze_face(crops[i], (0, 0, 100, 100), hs, fs, threads=4, features_mask=63)
            assert got[i]["headpose"].tobytes() == want["headpose"].tobytes() and got[i]["variance"].tobytes() == want["variance"].tobytes()
            assert np.array_equal(got[i]["tree_counts"], want["tree_counts"]) and np.array_equal(got[i]["n_votes"], want["n_votes"])
            assert np.array_equal(got[i]["ms_iters"], want["ms_iters"]) and np.abs(got[i]["ffd_f"] - want["ffd_f"]).max() <= 1e-3
    # a 3-plane model (features 0 and 2) and the same trees with MIN_MAX added at run time
    hp3, ffd3 = sm.write_model(tmp_path / "three", seed=22, channels=3, features=(0, 2))
    g3, o3 = crf.Model(hp3, ffd3, 15, 20), O.Model(hp3, ffd3, 15, 20)
    got = crf.Context(g3, 0).analyze_crops(crops[:4])
    for i in range(4):
        want = o3.analyze_face(crops[i], (0, 0, 100, 100), features_mask=0b101)
        assert got[i]["headpose"].tobytes() == want["headpose"].tobytes() and np.array_equal(got[i]["ffd"], want["ffd"])
    g3.set_features([0, 2, 3])
    got2 = crf.Context(g3, 0).analyze_crops(crops[:4])
    assert got2.tobytes() == got.tobytes()   # the trees never read the two added planes


def test_evaluate_on_explicit_patches_and_eval_test(O, crf, gpu, synth_models, synth_dirs):
    """Forest<S>::evaluateMT per sample and ImageSample::evalTest: the dense-grid results picked at arbitrary patches, the
    integer mean-difference formula, and the same through contexts that hold a single forest."""
    gm, om = synth_models
    hp, ffd = synth_dirs
    rng = np.random.default_rng(9)
    planes = rng.integers(0, 256, (38, 140, 125), dtype=np.uint8)
    ctx = crf.Context(gm, 0)
    dense = ctx.stage_eval_forest(planes, 1)                       # [x * ny + y][tree]
    ny = 140 - 31
    xy = np.stack([rng.integers(0, 125 - 31, 50), rng.integers(0, ny, 50)], 1)
    pick = dense[xy[:, 0] * ny + xy[:, 1]]
    assert np.array_equal(ctx.stage_eval_patches(planes, xy), pick)
    fi = rng.integers(0, 5, 20); ti = rng.integers(0, 20, 20)
    dense_f = ctx.stage_eval_forest(planes, 1, fi, ti)
    assert np.array_equal(ctx.stage_eval_patches(planes, xy, fi, ti), dense_f[xy[:, 0] * ny + xy[:, 1]])
    # single-forest contexts (Forest<S>::load on its own)
    c_hp = crf.Context(crf.Model(forest_dir=hp, kind="hp", ntrees=15), 0)
    assert np.array_equal(c_hp.stage_eval_patches(planes, xy), pick)
    r_full, r_hp = ctx.stage_headpose(planes, 4), c_hp.stage_headpose(planes, 4)
    assert r_hp["headpose"] == r_full["headpose"] and r_hp["variance"] == r_full["variance"] and len(r_hp["forest_idx"]) == 0
    c_mp = crf.Context(crf.Model(forest_dir=str(Path(ffd) / "forest_2"), kind="mp", ntrees=20), 0)
    t2 = np.arange(20)
    assert np.array_equal(c_mp.stage_eval_patches(planes, xy, np.zeros(20, int), t2), ctx.stage_eval_patches(planes, xy, np.full(20, 2), t2))
    for bad in (lambda: c_hp.stage_eval_forest(planes, 3, fi, ti), lambda: c_mp.stage_headpose(planes, 4), lambda: c_hp.analyze_crops(np.zeros((1, 100, 100, 3), np.uint8)),
                lambda: ctx.stage_eval_patches(planes, [[100, 5]])):
        with pytest.raises(crf.CrfError):
            bad()
    # evalTest
    integ = np.zeros((38, 141, 126), np.int64)
    integ[:, 1:, 1:] = planes.astype(np.int64).cumsum(1).cumsum(2)
    tests, want = [], []
    for _ in range(500):
        c = int(rng.integers(0, 38)); px = int(rng.integers(0, 125 - 31)); py = int(rng.integers(0, 140 - 31))
        r = []
        for _k in range(2):
            w, h = int(rng.integers(1, 23)), int(rng.integers(1, 23))
            r.append((int(rng.integers(0, 31 - w)), int(rng.integers(0, 31 - h)), w, h))
        m = [int(integ[c, py + y + h, px + x + w] - integ[c, py + y, px + x + w] - integ[c, py + y + h, px + x] + integ[c, py + y, px + x]) // (w * h) for (x, y, w, h) in r]
        tests.append([c, *r[0], *r[1], px, py]); want.append(m[0] - m[1])
    assert np.array_equal(crf.Context(None, 0).stage_eval_tests(planes, tests), np.array(want, np.int32))
    # the !m_use_integral branch (cv::sum over the 8-bit rectangles, src/ImageSample.cpp:40-47) gives the same integers
    assert np.array_equal(crf.Context(None, 0).stage_eval_tests(planes, tests, use_integral=False), np.array(want, np.int32))


def test_multi_gpu_single_caller(crf, O, gpu, synth_models):
    """crf_multi_*: shards behind one caller.  On a one-GPU box the same device is listed three times, which exercises the
    sharding, the threads and the frame-boundary cuts exactly as three GPUs would; results must equal the single-context ones."""
    from face_alignment_cvpr_2012_b200 import workloads as wl
    gm, _ = synth_models
    ndev = crf.lib().crf_device_count()
    devices = list(range(ndev)) if ndev >= 2 else [0, 0, 0]
    mc = crf.MultiContext(gm, devices)
    assert mc.n_devices == len(devices)
    one = crf.Context(gm, 0)
    crops, _ = wl.make_crops(37, seed=12)
    assert mc.analyze_crops(crops).tobytes() == one.analyze_crops(crops).tobytes()
    assert mc.analyze_crops(crops[:2]).tobytes() == one.analyze_crops(crops[:2]).tobytes()      # fewer faces than shards
    assert mc.analyze_crops(crops, headpose_only=True).tobytes() == one.analyze_crops(crops, headpose_only=True).tobytes()
    frames, boxes, iob, _ = wl.make_frames(5, rows=360, cols=480, faces_per_frame=3, seed=3, wmin=64, wmax=110)
    assert mc.analyze_batch(frames, boxes, iob).tobytes() == one.analyze_batch(frames, boxes, iob).tobytes()
    bad = boxes.copy(); bad[len(bad) - 1] = (470, 10, 100, 100)   # the last shard's last box leaves its frame: the error names the shard
    with pytest.raises(crf.CrfError) as e:
        mc.analyze_batch(frames, bad, iob)
    assert e.value.code == -1 and "shard" in str(e.value)
    assert crf.MultiContext(gm).n_devices == ndev


def test_fused_and_banded_gabor_kernels_agree(crf, O, gpu, synth_models, monkeypatch):
    """The fused per-scale Gabor kernels (CRF_GABOR_FUSED=1: rolling row pass, magnitudes in an L2-resident per-CTA scratch,
    quantisation in the same kernel; measured slower, kept as a variant) and both quantisation kernels against the default banded path and the oracle:
    identical planes on ragged heights (1 .. 33 bands, partial last band, W = 124 / 125) and identical records in a batch."""
    import cv2
    from face_alignment_cvpr_2012_b200 import workloads as wl
    gm, _ = synth_models
    rng = np.random.default_rng(0)
    imgs = [cv2.GaussianBlur(rng.integers(0, 256, (H, W), dtype=np.uint8), (0, 0), 1.3) for H, W in [(125, 125), (148, 124), (32, 125), (47, 125), (48, 125), (49, 124), (521, 125), (200, 125)]]
    imgs.append(np.zeros((125, 125), np.uint8)); imgs.append(rng.integers(0, 256, (125, 125), dtype=np.uint8))
    banded = crf.Context(None, 0)
    monkeypatch.setenv("CRF_GABOR_FUSED", "1")
    fused = crf.Context(None, 0)
    monkeypatch.delenv("CRF_GABOR_FUSED")
    monkeypatch.setenv("CRF_GABOR_QUANT_OLD", "1")
    old_quant = crf.Context(None, 0)
    monkeypatch.delenv("CRF_GABOR_QUANT_OLD")
    monkeypatch.setenv("CRF_GABOR_BAND", "32" if os.environ.get("CRF_GABOR_BAND", "") != "32" else "16")   # the band height that is NOT the default
    other_band = crf.Context(None, 0)
    monkeypatch.delenv("CRF_GABOR_BAND")
    for img in imgs:
        pf, jf = fused.stage_channels(img)
        pb, jb = banded.stage_channels(img)
        pq, jq = old_quant.stage_channels(img)
        po, jo = other_band.stage_channels(img)
        assert np.array_equal(pf, pb) and np.array_equal(jf, jb) and np.array_equal(pq, pb) and np.array_equal(jq, jb), img.shape
        assert np.array_equal(po, pb) and np.array_equal(jo, jb), img.shape
        op, oi = O.channels(img)
        assert np.array_equal(pf, op) and np.array_equal(jf, oi.astype(np.uint32)), img.shape
    crops, _ = wl.make_crops(700, seed=8)     # more faces than resident CTAs: every persistent CTA takes several items
    a = crf.Context(gm, 0).analyze_crops(crops)
    monkeypatch.setenv("CRF_GABOR_FUSED", "1")
    b = crf.Context(gm, 0).analyze_crops(crops)
    assert a.tobytes() == b.tobytes()
