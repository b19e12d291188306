"""The drop-in made literal: the compiled pybind11 module `PyProjectD` (projectd_core_b200/csrc/pyprojectd_module.cpp, on the
C ABI of libpd_b200.so) under the reference's UNMODIFIED Python env (pyprojectd/projectd_env.py, a git-ignored copy made by
`make -C oracle content`), compared step by step with the oracle driven through the same call sequence and with
BatchedProjectDEnv(num_envs=1)."""
import importlib
import math
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "projectd_core_b200")
REF_PY = os.path.join(ROOT, "oracle", "_ref", "pyprojectd")


def _import_module():
    if PKG not in sys.path:
        sys.path.insert(0, PKG)
    return importlib.import_module("PyProjectD")


def test_module_surface_matches_reference():
    """Every name the reference's PYBIND11_MODULE defines (PyProjectD.cpp:515-641) exists with the same attribute sets."""
    pd = _import_module()
    for fn in ("setSeed setLogFile clearLogFile writeLog createSimulator destroySimulator stepSimulator loadTrack unloadTrack addCar removeCar "
               "teleportCarToLocation teleportCarToPits teleportCarToSpline teleportCarByMode setCarAutoTeleport setCarControls setCarAssists getCarState "
               "setCarRawTune setCarTune setScoringVar getScoringVar launchPlaygroundInOwnThread initPlayground shutPlayground shutAll tickPlayground "
               "isPlaygroundInitialized isPlaygroundExited moveWindow resizeWindow setRenderHz setActiveSimulator setActiveCar getActiveSimulator getActiveCar").split():
        assert callable(getattr(pd, fn)), fn
    c = pd.CarControls()
    assert (c.steer, c.clutch, c.brake, c.handBrake, c.gas, c.isShifterSupported, c.requestedGearIndex, c.gearUp, c.gearDn) == (0, 0, 0, 0, 0, 1, -1, 0, 0)
    c.steer = 0.5; c.requestedGearIndex = 2
    st = pd.CarState()
    for name in ("carId simId timestamp controls collisionFlag outOfTrackFlag trackPointId lastTrackPointTimestamp trackLocation bodyVsTrack velocityVsTrack "
                 "engineRPM speedMS gear gearGrinding bodyMatrix bodyPos bodyEuler accG velocity localVelocity angularVelocity localAngularVelocity hubMatrix "
                 "tyreContacts tyreLoad tyreAngularSpeed tyreSlipRatio tyreNdSlip probes lookAhead stepReward totalReward").split():
        assert hasattr(st, name), name
    assert len(st.probes) == 10 and len(st.lookAhead) == 5 and len(st.hubMatrix) == 4 and hasattr(st.bodyMatrix, "M43") and hasattr(st.localVelocity, "z")
    with pytest.raises(AttributeError):
        st.gear = 3                       # read-only, as in the reference's bindings
    # reference error behaviour: bad ids are silent no-ops, nothing raises (PyProjectD.cpp:100-136)
    pd.stepSimulator(12345, 1 / 333.0); pd.getCarState(12345, 0, st); pd.setCarTune(12345, 0, "FRONT_BIAS", 55.0)
    assert pd.addCar(12345, "ks_toyota_ae86_drift") == -1 and pd.getScoringVar(12345, 0, "TravelBonus") == 0.0


def _reference_env_class(base_path):
    if not os.path.exists(os.path.join(REF_PY, "projectd_env.py")):
        pytest.skip("oracle/_ref/pyprojectd (copy of the reference's env) absent: run `make -C oracle content` where /root/reference exists")
    _import_module()
    if REF_PY not in sys.path:
        sys.path.insert(0, REF_PY)
    if not hasattr(os, "add_dll_directory"):
        os.add_dll_directory = lambda p: None       # projectd_env.py:12 is Windows-only; the file itself stays untouched
    mod = importlib.import_module("projectd_env")
    mod.base_dir = base_path                         # module global read by ProjectDEnv.__init__ (projectd_env.py:100-103)
    return mod.ProjectDEnv


@pytest.mark.gpu
def test_unmodified_reference_env_runs_on_the_module(oracle):
    """pyprojectd/projectd_env.py, byte for byte the reference's file, on the compiled PyProjectD module: reset + 600 steps of a
    scripted policy.  Step by step against (a) the oracle driven through the same PyProjectD call sequence (free running from
    the grid: bound grows with time; 1e-2 over the first 250 steps, 0.1 for the tyres' ndSlip during the wheel-spinning launch) and (b) BatchedProjectDEnv(num_envs=1) (same kernels:
    1e-5), including reward and terminate."""
    import torch
    Env = _reference_env_class(oracle.BASE_PATH)
    env = Env()
    from projectd_core_b200.env import BatchedProjectDEnv
    benv = BatchedProjectDEnv(oracle.BASE_PATH, num_envs=1, device=0, autoreset_mode=1)
    r = oracle.RefSim()                              # configured like projectd_env.py:118-136
    lay = oracle.Layout()

    def ref_obs():
        st = np.zeros(664, np.uint8); r.L.pdref_get_car_state(r.h, st.ctypes.data)
        from projectd_core_b200.pyprojectd import CAR_STATE_DTYPE
        s = np.frombuffer(st, dtype=CAR_STATE_DTYPE)[0]
        o = np.concatenate([s["localVelocity"], s["localAngularVelocity"], s["tyreNdSlip"], [s["bodyVsTrack"], s["velocityVsTrack"]], s["lookAhead"], s["probes"][:7]])
        return o.astype(np.float32), float(s["stepReward"]), int(s["collisionFlag"]), int(s["outOfTrackFlag"])

    obs0 = env.reset()
    bobs0 = benv.reset().cpu().numpy()[0]
    r.L.pdref_teleport_mode(r.h, 0); r.set_controls(steer=0.0, gas=0.55); r.step()          # reset(): teleport + step(action 0) -> gas = linscale(0) = 0.55
    assert obs0.shape == (24,) and obs0.dtype == np.float32
    assert np.abs(obs0 - bobs0).max() <= 1e-5 and np.abs(obs0 - ref_obs()[0]).max() <= 1e-4
    worst_b = worst_o = worst_nd = 0.0; trace = []
    for t in range(600):
        a = np.array([0.35 * math.sin(t / 90.0), min(1.0, -0.2 + t / 200.0)], np.float32)
        obs, rew, term, trunc, _ = env.step(a)
        bo, br, bt, _, _ = benv.step(torch.from_numpy(a[None, :]).cuda())
        bo = bo.cpu().numpy()[0]; br = float(br[0]); bt = bool(bt[0])
        r.set_controls(steer=float(a[0]), gas=float(0.1 + 0.9 * (a[1] + 1) * 0.5)); r.step()
        ro, rr, rc, roff = ref_obs()
        sc = np.maximum(np.abs(obs), 1.0)
        worst_b = max(worst_b, float((np.abs(obs - bo) / sc).max()), abs(rew - br))
        assert term == bt, t
        eo = (np.abs(obs - ro) / sc)
        trace.append((t, float(eo.max()), int(eo.argmax())))
        if t < 250:
            nd = eo[6:10].copy(); eo[6:10] = 0.0         # tyreNdSlip: slip of spinning tyres during the launch, chaotic -> own bound
            worst_o = max(worst_o, float(eo.max())); worst_nd = max(worst_nd, float(nd.max()))
            assert abs(rew - rr) <= 1e-3 and term == bool(rc or roff), t
        if term:
            break
    assert worst_b <= 1e-5, worst_b
    print("drop-in vs oracle, free running: (step, worst relative obs error, obs index) every 25 steps:", trace[::25])
    assert worst_o <= 1e-2 and worst_nd <= 0.1, (worst_o, worst_nd, trace[::25])
    assert env.dstate.speedMS > 2.0, "the scripted policy must get the car moving"
    env.close(); benv.close()


@pytest.mark.gpu
def test_reference_env_with_the_supra(oracle):
    """The reference's env file names a second car (`#car_model = 'ks_toyota_supra_mkiv_drift'`, projectd_env.py:25: double
    wishbones all round, two turbochargers).  A subclass that only switches `car_model` runs on the compiled module; its
    observations follow the oracle driven through the same call sequence (free running; no `car_tunes` entry exists for this
    car, so none are applied on either side)."""
    Env = _reference_env_class(oracle.BASE_PATH)

    class SupraEnv(Env):
        car_model = "ks_toyota_supra_mkiv_drift"

    env = SupraEnv()
    r = oracle.RefSim(car="ks_toyota_supra_mkiv_drift", env_setup=False)
    r.L.pdref_teleport_mode(r.h, 0); r.L.pdref_set_auto_teleport(r.h, 0, 0, 0); r.L.pdref_set_assists(r.h, 1, 1, 1)
    for k, v in oracle.ENV_SCORING.items():
        r.L.pdref_set_scoring_var(r.h, k.encode(), v)

    def ref_obs():
        st = np.zeros(664, np.uint8); r.L.pdref_get_car_state(r.h, st.ctypes.data)
        from projectd_core_b200.pyprojectd import CAR_STATE_DTYPE
        s = np.frombuffer(st, dtype=CAR_STATE_DTYPE)[0]
        o = np.concatenate([s["localVelocity"], s["localAngularVelocity"], s["tyreNdSlip"], [s["bodyVsTrack"], s["velocityVsTrack"]], s["lookAhead"], s["probes"][:7]])
        return o.astype(np.float32), float(s["stepReward"])

    obs0 = env.reset()
    r.L.pdref_teleport_mode(r.h, 0); r.set_controls(steer=0.0, gas=0.55); r.step()
    assert np.abs(obs0 - ref_obs()[0]).max() <= 1e-4
    worst = 0.0
    for t in range(650):      # the auto-shifter leaves neutral after ~1 s, the car is rolling at ~2 m/s after 1.5 s
        a = np.array([0.3 * math.sin(t / 80.0), min(1.0, -0.4 + t / 150.0)], np.float32)
        obs, rew, term, trunc, _ = env.step(a)
        r.set_controls(steer=float(a[0]), gas=float(0.1 + 0.9 * (a[1] + 1) * 0.5)); r.step()
        ro, rr = ref_obs()
        if t < 250:
            e = np.abs(obs - ro) / np.maximum(np.abs(obs), 1.0); e[6:10] = 0.0          # tyreNdSlip of the spinning rear tyres: chaotic, not bounded here
            worst = max(worst, float(e.max()))
            assert abs(rew - rr) <= 1e-3, t
        assert not term, t
    assert worst <= 1e-2, worst
    assert env.dstate.speedMS > 2.0 and env.dstate.engineRPM > 1500.0
    env.close()
