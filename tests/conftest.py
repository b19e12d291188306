import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_cuda():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    """The compiled reference (oracle/_ref/libpdref.so + its content copy)."""
    import pdref
    if not pdref.available():
        if os.path.isdir("/root/reference/src/ProjectD"):
            subprocess.check_call(["make", "-j8"], cwd=os.path.join(ROOT, "oracle"))
            subprocess.check_call(["make", "content"], cwd=os.path.join(ROOT, "oracle"))
        else:
            pytest.skip("oracle/_ref not built and /root/reference absent")
    return pdref


@pytest.fixture(scope="session")
def hostsim():
    """ctypes handle of tests/hostsim/libpdhostsim.so (device functions compiled for the host, debugging aid)."""
    import ctypes
    path = os.path.join(ROOT, "tests", "hostsim", "libpdhostsim.so")
    if not os.path.exists(path):
        subprocess.check_call(["make"], cwd=os.path.join(ROOT, "tests", "hostsim"))
    H = ctypes.CDLL(path)
    vp, cp, f, i, d = ctypes.c_void_p, ctypes.c_char_p, ctypes.c_float, ctypes.c_int, ctypes.c_double
    H.hs_create.restype = vp; H.hs_create.argtypes = [cp, cp, cp]
    H.hs_destroy.argtypes = [vp]
    H.hs_set_tune.argtypes = [vp, cp, f]
    H.hs_set_scoring_var.argtypes = [vp, cp, f]
    H.hs_set_assists.argtypes = [vp, i, i, i]
    H.hs_get_params.argtypes = [vp, vp]
    H.hs_get_track_info.argtypes = [vp, vp]
    H.hs_get_spline_nodes.argtypes = [vp, vp, vp]
    H.hs_tick.argtypes = [vp, vp, f, d]
    H.hs_tick_quad.argtypes = [vp, vp, f, d]
    H.hs_probe_compare.argtypes = [vp, i, vp, vp, vp]
    H.hs_params_bytes.restype = i
    H.hs_offset_autoshift_rpm.restype = i
    H.hs_teleport_point.argtypes = [vp, vp, i, d]
    H.hs_point_id_at_distance.argtypes = [vp, f]
    H.hs_raycast.argtypes = [vp, i, vp, vp]
    H.hs_sctm_solve.argtypes = [vp, i, i, vp, vp]
    return H


@pytest.fixture(scope="session")
def golden():
    """Committed golden vectors (tests/golden/demo_golden.npz, written by tests/golden/make_golden.py)."""
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "demo_golden.npz"))


@pytest.fixture(scope="session")
def content_base():
    """Directory holding cfg/ + content/ of the reference (a copy under oracle/_ref/base, made by `make -C oracle content`)."""
    import pdref
    if not os.path.isdir(pdref.BASE_PATH):
        if os.path.isdir("/root/reference/src/ProjectD"):
            subprocess.check_call(["make", "content"], cwd=os.path.join(ROOT, "oracle"))
        else:
            pytest.skip("no content copy (oracle/_ref/base) and /root/reference absent")
    return pdref.BASE_PATH


@pytest.fixture(scope="session")
def hostsim_env(hostsim, content_base):
    import pdref
    h = hostsim.hs_create(content_base.encode(), b"driftplayground", b"ks_toyota_ae86_drift")
    assert h
    hostsim.hs_set_assists(h, 1, 1, 1)
    for k, v in pdref.ENV_TUNES.items():
        hostsim.hs_set_tune(h, k.encode(), v)
    for k, v in pdref.ENV_SCORING.items():
        hostsim.hs_set_scoring_var(h, k.encode(), v)
    return h
