"""CPU: the C-ABI library loads, exports every symbol include/crf_b200.h declares, and refuses to run without a GPU."""
import ctypes as C
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


def _declared():
    hdr = (ROOT / "include" / "crf_b200.h").read_text()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(crf_[a-z0-9_]+)\s*\(", hdr)))


def test_every_declared_symbol_is_exported(crf):
    from face_alignment_cvpr_2012_b200 import capi
    L = C.CDLL(str(capi.LIB_PATH))
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), n
    assert sorted(capi.EXPORTS) == names


def test_options_default_are_the_reference_defaults(crf):
    o = crf.Options()
    crf.lib().crf_options_default(C.byref(o))
    assert (o.hp_stride, o.ffd_stride, o.ffd_min_samples, o.ms_kernel_size, o.ms_max_iterations) == (4, 3, 2, 10, 7)
    assert abs(o.ffd_min_pf - 0.25) < 1e-9 and abs(o.ms_stopping_criteria - 0.05) < 1e-7 and o.ffd_max_variance == 25.0


def test_no_cpu_fallback(crf, synth_models):
    if crf.lib().crf_device_count() > 0:
        pytest.skip("a GPU is present")
    gm, _ = synth_models
    with pytest.raises(crf.CrfError) as e:
        crf.Context(gm, 0)
    assert e.value.code == -4 and "no CPU fallback" in str(e.value)
    ff = crf.FaceForest(model=gm)  # reference behaviour: constructor prints, is_inizialized stays false, use asserts
    assert not ff.is_inizialized
    with pytest.raises(AssertionError):
        ff.analyzeFace(None, (0, 0, 10, 10))


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under the package may import, link or call it."""
    bad = re.compile(r"(^\s*(from|import)\s+oracle\b|libcrf_oracle|\borc_[a-z_]+\s*\(|oracle/crf_oracle|oracle\.oracle)", re.M)
    for p in (ROOT / "face_alignment_cvpr_2012_b200").rglob("*"):
        if p.suffix in (".py", ".cu", ".cuh", ".cc", ".h") or p.name == "Makefile":
            m = bad.search(p.read_text())
            assert m is None, (p, m.group(0))


def test_cpp_compat_header_builds_and_mirrors_error_behaviour(crf, synth_dirs, tmp_path):
    """include/crf_b200_compat.hpp (the reference-side binding of INTEGRATION.md) compiles with g++ and behaves like
    the reference's FaceForest when the forests are missing or, here, when there is no GPU."""
    import subprocess
    from face_alignment_cvpr_2012_b200 import capi
    exe = tmp_path / "compat_smoke"
    subprocess.run(["g++", "-std=c++17", "-O1", "-I", str(ROOT / "include"), str(ROOT / "tests" / "cpp" / "compat_smoke.cc"), "-o", str(exe),
                    "-L", str(capi.LIB_PATH.parent), "-lcrf_b200", f"-Wl,-rpath,{capi.LIB_PATH.parent}"], check=True)
    hp, ffd = synth_dirs
    r = subprocess.run([str(exe), hp, ffd], capture_output=True, text=True)
    assert r.returncode == 0 and "compat_smoke ok" in r.stdout, r.stdout + r.stderr


def test_detectface_box_postprocessing(crf):
    """FaceForest::detectFace's enlargement (src/FaceForest.cpp:147-157) and intersect (src/face_utils.cpp:325-347)."""
    assert crf.enlarge_detections([(100, 50, 100, 100)], 480, 640) == [(95, 50, 110, 130)]
    assert crf.enlarge_detections([(2, 400, 101, 101)], 480, 640) == [(0, 400, 108, 80)]       # clipped left and bottom
    assert crf.enlarge_detections([(600, 10, 60, 60)], 480, 640) == [(597, 10, 43, 78)]       # clipped right
    assert crf.intersect((700, 0, 10, 10), (0, 0, 640, 480)) == (0, 0, 0, 0)


def test_cpp_eval_ffd_driver_builds_and_fails_like_the_reference(crf, tmp_path):
    """examples/eval_ffd.cpp compiles against the compat header; without its config files it exits with EXIT_FAILURE after the
    reference's "Default ForestParam initialization" message (src/face_utils.cpp:128-139, src/eval_ffd.cpp:139-143)."""
    import subprocess
    from face_alignment_cvpr_2012_b200 import capi
    exe = tmp_path / "eval_ffd"
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I", str(ROOT / "include"), str(ROOT / "examples" / "eval_ffd.cpp"), "-o", str(exe),
                    f"-L{capi.LIB_PATH.parent}", "-lcrf_b200", f"-Wl,-rpath,{capi.LIB_PATH.parent}"], check=True)
    r = subprocess.run([str(exe), str(tmp_path / "missing_ffd.txt"), str(tmp_path / "missing_hp.txt")], capture_output=True, text=True, timeout=60)
    assert r.returncode == 1 and "Default ForestParam initialization" in r.stdout
