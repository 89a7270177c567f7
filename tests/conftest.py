import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = Path(__file__).resolve().parent / "golden"
REF_DATA = Path("/root/reference/data")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # Contexts created with default options run MeanShift in the reference's exact summation order under the tests, so that
    # iteration counts and integer landmarks compare bit for bit; the tolerance mode (the library's default) is requested
    # explicitly where it is tested (ms_mode="fast").
    os.environ.setdefault("CRF_MS_MODE", "exact")


@pytest.fixture(scope="session")
def native():
    """Builds libcrf_b200.so and the oracle if they are stale."""
    import __graft_entry__ as g
    g.build()
    return True


@pytest.fixture(scope="session")
def O(native):
    from oracle import oracle
    return oracle


@pytest.fixture(scope="session")
def crf(native):
    import face_alignment_cvpr_2012_b200
    return face_alignment_cvpr_2012_b200


@pytest.fixture(scope="session")
def synth_dirs(tmp_path_factory):
    from face_alignment_cvpr_2012_b200 import synthetic_model as sm
    d = tmp_path_factory.mktemp("synth_model")
    return sm.write_model(d, seed=5)


@pytest.fixture(scope="session")
def synth_models(crf, O, synth_dirs):
    hp, ffd = synth_dirs
    return crf.Model(hp, ffd, 15, 20), O.Model(hp, ffd, 15, 20)


@pytest.fixture(scope="session")
def staged_models(crf, O):
    from face_alignment_cvpr_2012_b200 import workloads as wl
    p = wl.staged_model_path()
    if p is None:
        pytest.skip("staged/model.crfb200 not present (run tools/stage_data.py where /root/reference exists)")
    return crf.Model(packed=str(p)), O.Model(packed=str(p))


@pytest.fixture(scope="session")
def cv2_golden():
    return np.load(GOLDEN / "cv2_stages.npz")


@pytest.fixture(scope="session")
def lfw_golden():
    return np.load(GOLDEN / "lfw_e2e.npz")


@pytest.fixture(scope="session")
def lfw_faces():
    from face_alignment_cvpr_2012_b200 import workloads as wl
    f = wl.load_lfw()
    if not f:
        pytest.skip("staged/imgs not present")
    return f
