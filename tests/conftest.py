import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "flou.jl_b200"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu)")


def _has_gpu():
    try:
        import flou_b200
        return flou_b200.device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # `-m gpu` on a box without a GPU must fail loudly, not skip: a silent pass would hide
    # a missing CUDA path.  `-m "not gpu"` never touches these tests.
    pass


@pytest.fixture(scope="session")
def gpu():
    import flou_b200
    n = flou_b200.device_count()
    if n <= 0:
        pytest.fail("no CUDA device visible: the B200 path has no CPU fallback")
    return n


@pytest.fixture(autouse=True)
def _x_trace_array_like_config4(monkeypatch):
    """The library keeps the x-face trace array only for states beyond 48 MB (config 4); the small
    meshes of this suite would all take the other branch.  Tests run the config-4 branch unless they
    choose (`xtrace` fixture, test_stage_without_x_trace_array)."""
    if "FLOU_B200_XTRACE" not in os.environ:
        monkeypatch.setenv("FLOU_B200_XTRACE", "1")


@pytest.fixture(params=["array", "none"])
def xtrace(request, monkeypatch):
    """Both branches of the x-face traces of collocated nodes: written by the element kernel into an
    array (large states) or read from u by the face kernel (states that fit in L2)."""
    monkeypatch.setenv("FLOU_B200_XTRACE", "1" if request.param == "array" else "0")
    return request.param
