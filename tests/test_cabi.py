"""The drop-in boundary: libpd_b200.so loads without a GPU, exports every symbol include/pd_batch.h declares, and
fails LOUDLY (no CPU fallback) when no CUDA device is present."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "pd_batch.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pd_[a-z0-9_]+)\s*\(", text)))


def _has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_library_exports_every_declared_symbol():
    from projectd_core_b200 import load_library, lib_path
    assert os.path.exists(lib_path()), "build first: python -c 'import __graft_entry__ as g; g.build()'"
    L = load_library()
    names = _declared_symbols()
    assert len(names) >= 35, names
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing


def test_header_is_plain_c():
    """No torch / C++ types in the ABI: the header compiles as C."""
    import subprocess, tempfile
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "t.c")
        open(src, "w").write('#include "pd_batch.h"\nint main(void){ return pd_state_words == 0; }\n')
        subprocess.check_call(["gcc", "-std=c99", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), src])


def test_layout_constants_agree():
    from projectd_core_b200 import load_library
    import pdref
    L = load_library()
    assert L.pd_state_words() == pdref.Layout().words == 664
    assert L.pd_obs_dim() == 24
    assert L.pd_car_state_bytes() == 664
    from projectd_core_b200.pyprojectd import CAR_STATE_DTYPE
    assert CAR_STATE_DTYPE.itemsize == 664


@pytest.mark.skipif(_has_cuda(), reason="checks the no-GPU behaviour")
def test_create_fails_loudly_without_gpu(content_base):
    from projectd_core_b200 import Batch, PdError, load_library
    with pytest.raises(PdError) as ei:
        Batch(content_base, n_envs=4)
    assert "no CPU path" in str(ei.value) or "CUDA" in str(ei.value)
    L = load_library()
    h = ctypes.c_void_p()
    rc = L.pd_create(content_base.encode(), b"driftplayground", b"ks_toyota_ae86_drift", 4, 0, ctypes.byref(h))
    assert rc != 0 and not h.value
    assert L.pd_last_error(None)
    # bad arguments are reported, never a crash
    assert L.pd_create(None, b"x", b"y", 1, 0, ctypes.byref(h)) != 0
    assert L.pd_create(b"/nonexistent", b"driftplayground", b"ks_toyota_ae86_drift", 1, 0, ctypes.byref(h)) != 0
    assert L.pd_step(None, ctypes.c_float(0.003), 1) != 0


@pytest.mark.skipif(_has_cuda(), reason="checks the no-GPU behaviour")
def test_pyprojectd_mirror_error_behaviour(content_base):
    """Reference semantics (PyProjectD.cpp:100-136): nothing raises, -1 sentinel, silent no-op on bad ids."""
    from projectd_core_b200 import pyprojectd as pd
    sim = pd.createSimulator(content_base)
    assert sim >= 0
    pd.loadTrack(sim, "driftplayground")
    assert pd.addCar(sim, "ks_toyota_ae86_drift") == -1          # no GPU here -> failure sentinel, no exception
    assert pd.addCar(12345, "ks_toyota_ae86_drift") == -1
    st = pd.CarState(); ctl = pd.CarControls()
    pd.setCarControls(sim, 0, True, ctl); pd.stepSimulator(sim, 1 / 333.0); pd.getCarState(sim, 0, st)   # no-ops
    pd.teleportCarByMode(999, 0, 0); pd.setCarTune(999, 0, "FRONT_BIAS", 55.0)
    assert st.speedMS == 0 and isinstance(st.bodyPos, pd.vec3f) and len(st.probes) == 10 and len(st.hubMatrix) == 4
    pd.destroySimulator(sim); pd.destroySimulator(sim)
