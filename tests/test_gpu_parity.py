"""GPU parity tests proper: the CUDA path, called through the C ABI (libpd_b200.so), against the oracle
(the reference's own Car/Sim/Core sources + restated ODE) on identical states and inputs."""
import math
import os

import numpy as np
import pytest

from parity_util import arbitrate, compare_records, coverage_marks, drive_controls, drive_start_states, make_env_like, scripted_controls

pytestmark = pytest.mark.gpu
DT = 1.0 / 333.0


@pytest.fixture(scope="module")
def lay(oracle):
    return oracle.Layout()


def _batch(oracle, n, **kw):
    from projectd_core_b200 import Batch
    return make_env_like(Batch(oracle.BASE_PATH, n_envs=n, device=0, **kw))


def test_library_is_native_cuda(oracle):
    """The product path is the CUDA library: it must be loaded and must launch kernels."""
    b = _batch(oracle, 4)
    b.teleport_spline(0.0); b.step(DT, 3); b.sync()
    assert b.launch_count() >= 5
    with open("/proc/self/maps") as f:
        assert "libpd_b200.so" in f.read()


def test_params_and_track_match_reference_init(oracle):
    b = _batch(oracle, 1)
    r = oracle.RefSim()
    ref = r.params_bytes()
    mine = b.params_bytes()[: len(ref)]
    assert np.array_equal(mine, ref), "car parameter block differs from the reference's own init"
    ti = np.zeros(12, np.uint32); r.L.pdref_get_track_info(r.h, ti.ctypes.data)
    info = b.track_info()
    assert info["nTris"] == int(ti[1]) and info["nFatPoints"] == int(ti[3]) and info["nSplineNodes"] == int(ti[4])
    assert info["computedTrackLength"] == float(ti[8:9].view(np.float32)[0])


def test_initial_and_teleport_state(oracle, lay):
    """create + teleportCarByMode(0) (projectd_env.py:123) gives the reference's state, field by field."""
    b = _batch(oracle, 3)
    b.teleport_spline(np.array([0.0, 0.37, 0.81], np.float32)); b.sync()
    for i, u in enumerate([0.0, 0.37, 0.81]):
        r = oracle.RefSim()
        r.teleport_spline(u)
        bad, worst = compare_records(lay, b.get_state(i), r.state(), tol=1e-6)
        assert not bad, bad[:10]


def test_raycast_bit_exact(oracle):
    rng = np.random.default_rng(3)
    fat = np.fromfile(oracle.BASE_PATH + "/content/tracks/driftplayground/spline.cache", dtype=np.float32).reshape(-1, 15)
    n = 50000
    idx = rng.integers(0, len(fat), n)
    rays = np.zeros((n, 7), np.float32)
    rays[:, 0:3] = fat[idx, 0:3] + rng.uniform(-12, 12, (n, 3)).astype(np.float32) * np.array([1, 0, 1], np.float32) + np.array([0, 2.5, 0], np.float32)
    rays[:, 4] = -1; rays[:, 6] = 3.0
    rays[n // 2:, 6] = 1000.0; rays[n // 2:, 1] += 10
    b = _batch(oracle, 1)
    mine = b.raycast(rays)
    r = oracle.RefSim()
    ref = np.zeros((n, 8), np.float32)
    r.L.pdref_raycast(r.h, n, rays.ctypes.data, ref.ctypes.data)
    assert np.array_equal(mine[:, 0], ref[:, 0]), "hit flags must match exactly"
    hit = ref[:, 0] == 1
    assert np.array_equal(mine[hit, 7], ref[hit, 7]), "surface ids must match exactly"
    assert np.abs(mine[hit, 1:7] - ref[hit, 1:7]).max() <= 1e-6


def test_spline_cache_known_answers(oracle):
    """The reference's own spline.cache pins ray-vs-trimesh hits (Track::computeFatPoints, Track.cpp:366-433)."""
    base = oracle.BASE_PATH + "/content/tracks/driftplayground/"
    slim = np.fromfile(base + "spline.bin", dtype=np.float32).reshape(-1, 5)
    fat = np.fromfile(base + "spline.cache", dtype=np.float32).reshape(-1, 15)
    rays = np.zeros((len(slim), 7), np.float32)
    rays[:, 0:3] = slim[:, 0:3] + np.array([0, 20, 0], np.float32); rays[:, 4] = -1; rays[:, 6] = 100
    out = _batch(oracle, 1).raycast(rays)
    assert out[:, 0].all()
    assert np.abs(out[:, 1:4] - fat[:, 0:3]).max() <= 1e-4


# every tick-kernel instance the dispatcher can pick (pd_batch.cu finish_create / launch_tick): the quad kernel at 2 / 4 / 8 cars
# per warp (helper quads at 2 and 4; 4 is what BASELINE configs[1] = 4096 envs runs, 8 serves 4097..8192 envs) and the
# thread-per-car kernel (what configs[2] = 65536 envs per GPU runs)
KERNELS = {
    "k_tick_quad<2>": {"PD_QUAD_MAX_ENVS": "20480", "PD_QUAD_CPW": "2"},
    "k_tick_quad<4>": {"PD_QUAD_MAX_ENVS": "20480", "PD_QUAD_CPW": "4"},
    "k_tick_quad<8>": {"PD_QUAD_MAX_ENVS": "20480", "PD_QUAD_CPW": "8"},
    "k_tick": {"PD_QUAD_MAX_ENVS": "0", "PD_SERIAL_WIDE": "0"},          # the 128-register instance: what batches beyond 37888 envs (the 65536-env bench) run
    "k_tick/255": {"PD_QUAD_MAX_ENVS": "0", "PD_SERIAL_WIDE": "1"},      # the 255-register instance: thread-per-car batches that fit one wave at 4 blocks per SM
}


def _select_kernel(monkeypatch, kernel):
    for k, v in KERNELS[kernel].items():
        monkeypatch.setenv(k, v)


def _single_tick_parity(oracle, lay, b, track, n_envs, ticks, want, car=None, resync=False):
    """Returns nothing; asserts the single-tick rule: ints exact; floats within 1e-4 (parity_util.compare_records); a record
    that leaves the rule is handed to the conditioning arbiter (parity_util.arbitrate: the oracle's own response to a one-ulp
    nudge of the body positions) and must be explained by it; such records must be rare (< 0.5 % of car-ticks)."""
    ckw = {"car": car} if car else {}
    starts = drive_start_states(oracle, lay, track, n_envs, **ckw)
    refs = [oracle.RefSim(track=track, **ckw) for _ in range(n_envs)]
    for r, (rec, tm, fr) in zip(refs, starts):
        r.set_state(rec); r.set_time(0.0)
    worst = 0.0; narb = 0; failures = []; seen = set()
    recs = [r.state() for r in refs]
    for t in range(ticks):
        for i, r in enumerate(refs):
            r.set_controls(**drive_controls(t, i, lay, recs[i]))
        before = [r.state() for r in refs]
        tb = refs[0].time()
        b.restore(np.stack(before, axis=1))
        b.set_time(tb)
        b.step(DT, 1)
        out = b.snapshot()
        for i, r in enumerate(refs):
            if resync:          # contact joints are not part of the record (a restored state has none alive, on either side): drives that
                r.set_state(before[i])      # meet walls re-enter the oracle through its record too, so both sides start every tick without joints
            r.step()
            ref = recs[i] = r.state()
            bad, w = compare_records(lay, out[:, i], ref, tol=1e-4)
            if bad:
                narb += 1
                left = arbitrate(oracle, lay, track, before[i], tb, ref, bad, **ckw)
                if left and len(failures) < 10:
                    failures.append((t, i, left[:4]))
            else:
                worst = max(worst, w)
            coverage_marks(lay, ref, seen)
    assert not failures, failures
    assert narb <= ticks * n_envs * 0.005, narb
    # the drives must really have reached the state space they were written for
    assert want <= seen, sorted(want - seen)
    return worst, narb


@pytest.mark.parametrize("kernel", list(KERNELS))
def test_single_tick_parity_identical_states(oracle, lay, kernel, monkeypatch):
    """For every tick of 64 scripted drives (parity_util.drive_controls: grid launches, handbrake turns / brake-to-lock at
    speed, reverse gear, standstill with the sleep counter, rev limiter, leaving the tarmac): load the oracle's state into the
    GPU batch, advance ONE tick on the GPU and in the oracle, compare all state fields.  Flags / FSM ints exact, floats within
    the rule of parity_util (north star: 1e-4 relative).  Run on EVERY tick-kernel instance, with blocks / tiles / helper
    quads filled."""
    _select_kernel(monkeypatch, kernel)
    b = _batch(oracle, 64)
    assert b.tick_kernel_instance() == kernel
    _single_tick_parity(oracle, lay, b, "driftplayground", 64, 500,
                        {"reverse", "handbrake", "locked_wheel_at_speed", "limiter", "sleeping", "gear>=3"})


OTHER_CARS = {  # the other four bundled cars: suspension pair as the kernel instances name it, topology id, turbochargers
    "ks_mazda_rx7_tuned": ("dwb,dwb", 3, 1), "ks_toyota_supra_mkiv_drift": ("dwb,dwb", 3, 2),
    "dthwsh_mazda_rx7_fc3s_sr20": ("strut,dwb", 1, 1), "gravygarage_street_ae86_readie": ("strut,dwb", 1, 1),
}


@pytest.mark.parametrize("car,kernel", [(c, k) for i, c in enumerate(OTHER_CARS) for k in ("k_tick_quad<4>" if i % 2 == 0 else "k_tick_quad<8>", "k_tick")])
def test_other_cars_params_and_single_tick_parity(oracle, lay, car, kernel, hostsim, monkeypatch):
    """double-wishbone suspensions (SuspensionDW.cpp:157-272: five distance joints per hub) and turbochargers
    (Turbo.cpp:11-40, Engine.cpp:368-384).  The loader's parameter block equals the reference's init byte for byte, the teleport
    state matches at 1e-6, and 48 scripted drives (the same script as the demo car's: handbrake, lock, reverse, limiter, sleep)
    hold the single-tick rule on the compile-time kernel instances of the car's suspension pair: thread per car, and 4 lanes per car
    (4 cars per warp for one car of each pair, 8 for the other)."""
    from projectd_core_b200 import Batch
    from parity_util import params_equal
    pair, topo, nturbo = OTHER_CARS[car]
    _select_kernel(monkeypatch, kernel)
    b = make_env_like(Batch(oracle.BASE_PATH, n_envs=48, device=0, car=car))
    want_inst = "k_tick<%s>" % pair if kernel == "k_tick" else kernel[:-1] + "," + pair + ">"
    assert b.tick_kernel_instance() == want_inst and b.topology() == topo
    r = oracle.RefSim(car=car)
    assert params_equal(b.params_bytes(), r.params_bytes(), hostsim), "car parameter block differs from the reference's own init"
    b.teleport_spline(np.full(48, 0.37, np.float32)); b.sync()
    r2 = oracle.RefSim(car=car); r2.teleport_spline(0.37)
    bad, worst = compare_records(lay, b.get_state(5), r2.state(), tol=1e-6)
    assert not bad, bad[:10]
    worst, narb = _single_tick_parity(oracle, lay, b, "driftplayground", 48, 300, {"reverse", "handbrake"}, car=car, resync=True)
    if nturbo:
        off = lay.fields["car.turboBoost"][0]
        assert b.snapshot()[off].view(np.float32).max() > 0.05, "the drives never built boost"


@pytest.mark.parametrize("kernel", ["k_tick_quad<4>", "k_tick"])
def test_single_tick_parity_offtrack_surfaces(oracle, lay, kernel, monkeypatch):
    """The same on yamanashi_short, whose grass (SIN_HEIGHT 0.03 / SIN_LENGTH 0.5) and sand (DAMPING 0.1, DIRT_ADDITIVE 1,
    SIN_HEIGHT 0.04) exercise the surface branches of Tyre::step / addGroundContact / stepDirtyLevel (Tyre.cpp:558-602,
    TyreForces.cpp:212-233) that driftplayground's plain surfaces never reach."""
    _select_kernel(monkeypatch, kernel)
    from projectd_core_b200 import Batch
    b = make_env_like(Batch(oracle.BASE_PATH, track="yamanashi_short", n_envs=64, device=0))
    assert b.tick_kernel_instance() == kernel
    _single_tick_parity(oracle, lay, b, "yamanashi_short", 64, 450, {"dirty_tyre", "gear>=3"})


@pytest.mark.parametrize("kernel", list(KERNELS))
def test_free_running_trajectory_divergence_1s(oracle, lay, kernel, monkeypatch):
    """333 ticks (1 s) free running from the same start with the same controls: bounded divergence.
    The drive is a full-throttle first-gear launch of a drift car with sinusoidal steering, i.e. tyres at
    saturation, where round-off differences grow quickly; bound: 5 cm / 0.5 deg after 1 s (measured worst on
    B200: 1.6 cm), gear state identical."""
    _select_kernel(monkeypatch, kernel)
    n = 64
    b = _batch(oracle, n)
    assert b.tick_kernel_instance() == kernel
    refs = [oracle.RefSim() for _ in range(n)]
    us = np.array([(i % 8) / 8 + (i // 8) * 0.007 for i in range(n)], np.float32)
    for i, r in enumerate(refs):
        r.teleport_spline(float(us[i]))
    b.teleport_spline(us)
    ctl = np.zeros((n, 5), np.float32)
    for t in range(333):
        for i, r in enumerate(refs):
            steer, gas = scripted_controls(t, phase=0.5 * i)
            r.set_controls(steer=steer, gas=gas); r.step()
            ctl[i, 0] = steer; ctl[i, 4] = gas
        b.set_controls(ctl, None, smooth=True)
        b.step(DT, 1)
    out = b.snapshot()
    for i, r in enumerate(refs):
        ref = r.state()
        dp = [lay.get(out[:, i], "chassis." + k) - lay.get(ref, "chassis." + k) for k in ("px", "py", "pz")]
        assert math.sqrt(sum(d * d for d in dp)) <= 0.05, ("position diverged", i, dp)
        dq = [lay.get(out[:, i], "chassis." + k) - lay.get(ref, "chassis." + k) for k in ("qw", "qx", "qy", "qz")]
        assert 2 * math.sqrt(sum(d * d for d in dq)) <= math.radians(0.5), ("heading diverged", i, dq)
        assert lay.get(out[:, i], "car.currentGear") == lay.get(ref, "car.currentGear")


def test_shard_invariance(oracle):
    """N envs on one batch == the same envs split over two batches (env-id keyed RNG): bitwise."""
    n = 32
    full = _batch(oracle, n); full.set_seed(1234, 0)
    a = _batch(oracle, n // 2); a.set_seed(1234, 0)
    c = _batch(oracle, n // 2); c.set_seed(1234, n // 2)
    rng = np.random.default_rng(0)
    for bb in (full, a, c):
        bb.teleport_mode(2)
    for t in range(40):
        act = rng.uniform(-1, 1, (n, 2)).astype(np.float32)
        full.set_actions(act); a.set_actions(act[: n // 2]); c.set_actions(act[n // 2:])
        for bb in (full, a, c):
            bb.step(DT, 1)
    s = full.snapshot()
    assert np.array_equal(s[:, : n // 2], a.snapshot()) and np.array_equal(s[:, n // 2:], c.snapshot())


def test_obs_dlpack_and_env_step(oracle):
    import torch
    n = 256
    b = _batch(oracle, n); b.set_seed(7, 0)
    b.teleport_mode(0)
    obs = b.obs_tensor()
    assert obs.is_cuda and obs.shape == (n, 24) and obs.dtype == torch.float32
    act = torch.zeros((n, 2), device="cuda", dtype=torch.float32); act[:, 1] = 1.0
    rew = torch.zeros(n, device="cuda"); done = torch.zeros(n, device="cuda", dtype=torch.int32)
    for t in range(900):
        b.env_step(act, DT, None, rew, done)
    b.sync()
    host = b.obs_host()
    assert np.allclose(obs.cpu().numpy(), host)
    assert np.isfinite(host).all()
    assert np.abs(host[:, 0:3]).max(axis=1).mean() > 0.5, "cars should be moving under full throttle"
    st = b.env_stats()
    assert st[0] >= 0


def test_pyprojectd_mirror_matches_reference_car_state(oracle):
    """The reference's own call sequence (projectd_env.py:118-136,156-171) through the PyProjectD mirror; CarState
    (664-byte layout of Car/CarState.h) against the reference's getCarState after 700 free-running ticks
    (an API-level check: tolerances are those of a 2 s free run, the parity tests proper are above)."""
    from projectd_core_b200 import pyprojectd as pd
    sim = pd.createSimulator(oracle.BASE_PATH)
    assert sim >= 0
    pd.loadTrack(sim, "driftplayground")
    car = pd.addCar(sim, "ks_toyota_ae86_drift")
    assert car == 0
    pd.teleportCarByMode(sim, car, 0)
    pd.setCarAutoTeleport(sim, car, False, False, 0)
    pd.setCarAssists(sim, car, True, True, True)
    for k, v in oracle.ENV_TUNES.items():
        pd.setCarTune(sim, car, k, v)
    for k, v in oracle.ENV_SCORING.items():
        pd.setScoringVar(sim, car, k, v)
    pd.teleportCarToSpline(sim, car, 0.25)
    r = oracle.RefSim(); r.teleport_spline(0.25)
    ctl = pd.CarControls(); st = pd.CarState()
    for t in range(700):
        ctl.steer = 0.1 * math.sin(t / 80.0); ctl.gas = 0.8
        pd.setCarControls(sim, car, True, ctl)
        pd.stepSimulator(sim, DT)
        r.set_controls(steer=ctl.steer, gas=ctl.gas); r.step()
    pd.getCarState(sim, car, st)
    ref = np.zeros(664, np.uint8); r.L.pdref_get_car_state(r.h, ref.ctypes.data)
    ref = np.frombuffer(ref, dtype=pd.CAR_STATE_DTYPE)[0]
    assert st.gear == int(ref["gear"]) and st.trackPointId == int(ref["trackPointId"])
    assert st.collisionFlag == int(ref["collisionFlag"]) and st.outOfTrackFlag == int(ref["outOfTrackFlag"])
    assert float(ref["speedMS"]) > 1.0, "the drive must get the car moving"
    assert abs(st.speedMS - float(ref["speedMS"])) <= 0.05 + 2e-2 * float(ref["speedMS"])
    assert abs(st.engineRPM - float(ref["engineRPM"])) <= 3e-2 * float(ref["engineRPM"])
    assert np.allclose(list(st.bodyPos), ref["bodyPos"], atol=0.3)
    assert np.allclose(list(st.localVelocity), ref["localVelocity"], atol=0.3)
    assert np.allclose(st.probes, ref["probes"], atol=0.5) and np.allclose(st.lookAhead, ref["lookAhead"], atol=2e-2)
    assert np.allclose(st.tyreLoad, ref["tyreLoad"], rtol=0.1, atol=100.0)
    assert np.allclose([st.bodyMatrix.M41, st.bodyMatrix.M42, st.bodyMatrix.M43], ref["bodyMatrix"][12:15], atol=0.3)
    assert abs(st.timestamp - float(ref["timestamp"])) <= 1e-4
    pd.destroySimulator(sim)


def test_batched_env_reset_step(oracle):
    import torch
    from projectd_core_b200.env import BatchedProjectDEnv
    env = BatchedProjectDEnv(oracle.BASE_PATH, num_envs=1024, device=0, seed=3, teleport_mode=2, autoreset_mode=1)
    obs = env.reset()
    assert obs.shape == (1024, 24) and obs.is_cuda
    lo, hi = env.observation_bounds()
    total = torch.zeros(1024, device="cuda")
    for t in range(400):
        a = torch.rand((1024, 2), device="cuda") * 2 - 1
        obs, rew, term, trunc, _ = env.step(a)
        total += rew
    assert torch.isfinite(obs).all() and torch.isfinite(total).all()
    o = obs.cpu().numpy()
    # the reference's observation_space bounds are nominal (values are not clipped, tyreNdSlip can exceed 10);
    # bounded quantities must respect them: direction cosines, look-ahead angles, probe distances
    assert lo.shape == hi.shape == (24,)
    assert (o[:, 10:12] >= -1 - 1e-4).all() and (o[:, 10:12] <= 1 + 1e-4).all()
    assert (np.abs(o[:, 12:17]) <= math.pi + 1e-4).all()
    assert (o[:, 17:24] >= 0).all() and (o[:, 17:24] <= 55 + 1e-3).all()   # rays are cast to 1.1 x probe length (Track.cpp:497-562)
    st = env.episode_stats()
    assert st["nan"] == 0
    env.close()


def test_env_step_host_equals_device_path(oracle):
    """pd_env_step_host (host buffers) == pd_env_step (device buffers) on identically seeded batches."""
    import torch
    n = 64
    a = _batch(oracle, n); a.set_seed(5, 0); a.teleport_mode(2)
    c = _batch(oracle, n); c.set_seed(5, 0); c.teleport_mode(2)
    rng = np.random.default_rng(1)
    obs_h = np.zeros((n, 24), np.float32); rew_h = np.zeros(n, np.float32); done_h = np.zeros(n, np.int32)
    rew_d = torch.zeros(n, device="cuda"); done_d = torch.zeros(n, device="cuda", dtype=torch.int32)
    for t in range(60):
        act = rng.uniform(-1, 1, (n, 2)).astype(np.float32)
        a.env_step_host(act, DT, obs_h, rew_h, done_h)
        c.env_step(torch.from_numpy(act).cuda(), DT, None, rew_d, done_d)
    c.sync()
    assert np.array_equal(obs_h, c.obs_tensor().cpu().numpy())
    assert np.array_equal(rew_h, rew_d.cpu().numpy()) and np.array_equal(done_h, done_d.cpu().numpy())
    assert np.array_equal(a.snapshot(), c.snapshot())


@pytest.mark.parametrize("kernel", ["k_tick_quad", "k_tick"])
def test_autoreset_next_step_equals_same_step(oracle, kernel, monkeypatch):
    """PD_AUTORESET_NEXT_STEP (reset inside the next step's single kernel launch) must produce, one step later, exactly
    the reset observation PD_AUTORESET_SAME_STEP produces (teleport + zero-action tick, projectd_env.py:216-227), with
    the same done / reward on the finishing step and (0, 0) on the reset step."""
    import torch
    monkeypatch.setenv("PD_QUAD_MAX_ENVS", "20480" if kernel == "k_tick_quad" else "0"); monkeypatch.setenv("PD_SERIAL_WIDE", "0")
    n = 64
    a = _batch(oracle, n); a.set_seed(5, 0); a.teleport_spline(np.linspace(0, 0.9, n))
    c = _batch(oracle, n); c.set_seed(5, 0); c.teleport_spline(np.linspace(0, 0.9, n)); c.set_autoreset(1)
    assert a.tick_kernel() == kernel and c.tick_kernel() == kernel
    act = torch.zeros((n, 2), device="cuda"); act[:, 0] = torch.linspace(-1, 1, n, device="cuda"); act[:, 1] = 1.0
    ra = torch.zeros(n, device="cuda"); da = torch.zeros(n, device="cuda", dtype=torch.int32)
    rc = torch.zeros(n, device="cuda"); dc = torch.zeros(n, device="cuda", dtype=torch.int32)
    alive = np.ones(n, bool)            # envs whose two batches are still in lock step (no reset yet)
    pend = {}                           # env -> reset observation of the same-step batch
    checked = 0
    for t in range(1500):
        a.env_step(act, DT, None, ra, da); c.env_step(act, DT, None, rc, dc)
        a.sync(); c.sync()                      # the batches run on their own streams
        oa = a.obs_tensor().cpu().numpy(); oc = c.obs_tensor().cpu().numpy()
        da_h, dc_h, ra_h, rc_h = da.cpu().numpy(), dc.cpu().numpy(), ra.cpu().numpy(), rc.cpu().numpy()
        for e, want in list(pend.items()):     # the step after the finishing one
            assert dc_h[e] == 0 and rc_h[e] == 0.0
            assert np.array_equal(oc[e], want), "reset observation differs for env %d" % e
            checked += 1; del pend[e]
        assert np.array_equal(da_h[alive], dc_h[alive]) and np.array_equal(ra_h[alive], rc_h[alive])
        for e in np.nonzero(alive & (da_h != 0))[0]:
            pend[int(e)] = oa[e].copy(); alive[e] = False
        if not alive.any() and not pend:
            break
    assert checked >= n // 4, "too few episodes finished to check (%d)" % checked
    sa, sc = a.env_stats(), c.env_stats()
    assert sc[0] >= checked


@pytest.mark.parametrize("kernel", ["k_tick_quad", "k_tick"])
def test_collision_flag_matches_oracle(oracle, lay, golden, kernel, monkeypatch):
    """SURVEY.md A14 on the GPU: collisionFlag after one tick on an odd physics frame for the golden collision cases
    (whole car translated towards walls / into the ground) and for fresh random cases checked against the oracle live."""
    monkeypatch.setenv("PD_QUAD_MAX_ENVS", "20480" if kernel == "k_tick_quad" else "0"); monkeypatch.setenv("PD_SERIAL_WIDE", "0")
    flags = golden["coll_flag"]; n = len(flags)
    b = _batch(oracle, n)
    assert b.tick_kernel() == kernel
    b.restore(np.ascontiguousarray(golden["coll_before"].T)); b.set_time(float(golden["coll_time"][0]))
    b.step(DT, 1)
    out = b.snapshot()
    off = lay.fields["car.collisionFlag"][0]
    got = out[off].view(np.int32)
    assert np.array_equal(got, flags), np.nonzero(got != flags)[0][:10]
    # live cases: yawed cars near walls
    rng = np.random.default_rng(11)
    r = oracle.RefSim()
    recs, want = [], []
    bodies = ["chassis", "tank", "hub0", "strut0", "hub1", "strut1", "axle", "hub3"]
    for case in range(64):
        r.teleport_spline(float(rng.uniform(0, 1)))
        for t in range(30):
            r.set_controls(steer=float(rng.uniform(-1, 1)), gas=1.0); r.step()
        rec = r.state().copy()
        dx, dz = rng.uniform(-8, 8, 2)
        for bd in bodies:
            lay.set(rec, bd + ".px", lay.get(rec, bd + ".px") + dx); lay.set(rec, bd + ".pz", lay.get(rec, bd + ".pz") + dz)
        lay.set(rec, "car.physFrame", 1)
        r.set_state(rec); recs.append(r.state().copy()); r.step()
        want.append(lay.get(r.state(), "car.collisionFlag"))
    b2 = _batch(oracle, 64)
    b2.restore(np.ascontiguousarray(np.stack(recs).T)); b2.set_time(r.time() - DT)
    b2.step(DT, 1)
    got2 = b2.snapshot()[off].view(np.int32)
    assert np.array_equal(got2, np.array(want, np.int32))


def test_env_step_host_zero_copy_equals_device_path(oracle):
    """pd_env_step_host with page-locked host buffers and the one-launch step: the kernel reads the actions from and writes
    obs / reward / done to the caller's pinned memory directly; results must equal the device-buffer path bit for bit."""
    import torch
    n = 256
    a = _batch(oracle, n); a.set_seed(5, 0); a.teleport_mode(2); a.set_autoreset(1)
    c = _batch(oracle, n); c.set_seed(5, 0); c.teleport_mode(2); c.set_autoreset(1)
    h_act = torch.empty((n, 2), dtype=torch.float32).pin_memory(); h_obs = torch.zeros((n, 24), dtype=torch.float32).pin_memory()
    h_rew = torch.zeros(n, dtype=torch.float32).pin_memory(); h_done = torch.zeros(n, dtype=torch.int32).pin_memory()
    rew_d = torch.zeros(n, device="cuda"); done_d = torch.zeros(n, device="cuda", dtype=torch.int32)
    gen = torch.Generator(); gen.manual_seed(3)
    for t in range(80):
        h_act.copy_(torch.rand((n, 2), generator=gen) * 2 - 1)
        a.env_step_host(h_act, DT, h_obs, h_rew, h_done)
        c.env_step(h_act.cuda(), DT, None, rew_d, done_d); c.sync()
        assert torch.equal(h_obs, c.obs_tensor().cpu()) and torch.equal(h_rew, rew_d.cpu()) and torch.equal(h_done, done_d.cpu()), t
    assert np.array_equal(a.snapshot(), c.snapshot())


def test_synthetic_large_track_config4(oracle):
    """BASELINE configs[3]: a generated ~1 M-triangle circuit (20.8 km).  The vertical-ray column grid must agree with the BVH
    on it, and a batch must drive on it (finite state, cars moving, no false collisions on the open circuit)."""
    import torch
    from projectd_core_b200 import Batch
    n = 256
    b = make_env_like(Batch(oracle.BASE_PATH, n_envs=n, device=0, synthetic_tris=1000000))
    ti = b.track_info()
    assert 900000 <= ti["nTris"] <= 1100000 and ti["nFatPoints"] > 10000
    b.set_seed(3, 0); b.teleport_mode(2); b.set_autoreset(1); b.sync()
    st = b.snapshot()
    lay = oracle.Layout()
    px = st[lay.fields["chassis.px"][0]].view(np.float32); py = st[lay.fields["chassis.py"][0]].view(np.float32); pz = st[lay.fields["chassis.pz"][0]].view(np.float32)
    rays = np.zeros((n, 7), np.float32); rays[:, 0] = px; rays[:, 1] = py + 5; rays[:, 2] = pz; rays[:, 4] = -1; rays[:, 6] = 50
    down = b.raycast(rays)
    tilt = rays.copy(); tilt[:, 3] = 1e-6; tilt[:, 4] = -1.0                     # not exactly vertical -> BVH traversal
    tilt[:, 3:6] /= np.linalg.norm(tilt[:, 3:6], axis=1, keepdims=True)
    gen = b.raycast(tilt)
    assert down[:, 0].all() and np.array_equal(down[:, 0], gen[:, 0]) and np.array_equal(down[:, 7], gen[:, 7])
    assert np.abs(down[:, 1:4] - gen[:, 1:4]).max() < 1e-3
    act = torch.zeros((n, 2), device="cuda"); act[:, 1] = 1.0
    rew = torch.zeros(n, device="cuda"); done = torch.zeros(n, device="cuda", dtype=torch.int32)
    for t in range(900):
        b.env_step(act, DT, None, rew, done)
    b.sync()
    obs = b.obs_tensor().cpu().numpy()
    assert np.isfinite(obs).all() and np.abs(obs[:, 0:3]).max(axis=1).mean() > 0.5, "cars should be moving under full throttle"
    s = b.env_stats()
    assert s[3] == 0 and s[7] == 0, "no collisions / NaNs expected on the open generated circuit"


@pytest.mark.parametrize("n", [1, 5, 37])
def test_ragged_batch_sizes_match_full_batch(oracle, n):
    """Batch sizes that do not fill a block / a warp / a quad group: every env of an n-env batch must end in exactly the state
    the same env reaches inside a 64-env batch (same start point, same actions; collision frames included)."""
    import torch
    us = np.linspace(0.05, 0.95, 64).astype(np.float32)
    full = _batch(oracle, 64); full.set_seed(9, 0); full.teleport_spline(us); full.set_autoreset(1)
    part = _batch(oracle, n); part.set_seed(9, 0); part.teleport_spline(us[:n]); part.set_autoreset(1)
    act = torch.zeros((64, 2), device="cuda"); act[:, 0] = torch.linspace(-0.6, 0.6, 64, device="cuda"); act[:, 1] = 0.8
    rf = torch.zeros(64, device="cuda"); df = torch.zeros(64, device="cuda", dtype=torch.int32)
    rp = torch.zeros(n, device="cuda"); dp = torch.zeros(n, device="cuda", dtype=torch.int32)
    for t in range(300):
        full.env_step(act, DT, None, rf, df); part.env_step(act[:n].contiguous(), DT, None, rp, dp)
    full.sync(); part.sync()
    assert np.array_equal(full.snapshot()[:, :n], part.snapshot())
    assert torch.equal(rf[:n].cpu(), rp.cpu()) and torch.equal(df[:n].cpu(), dp.cpu())
    assert torch.equal(full.obs_tensor()[:n].cpu(), part.obs_tensor().cpu())


@pytest.mark.parametrize("kernel", ["k_tick_quad<4>", "k_tick"])
def test_car_state_and_obs_identical_state(oracle, lay, kernel, monkeypatch):
    """SURVEY.md A12 from IDENTICAL states: one tick from the oracle's state, then the 664-byte CarState (Car/CarState.h:11-56,
    Car::updateCarState Car.cpp:802-865) and the env's 24 observations (projectd_env.py:237-275) against the reference's own
    getCarState: ints exact, floats within 1e-4 of the field's scale (vectors against their norm; matrices: rotation entries
    against 1, the translation row against the tick's displacement scale), every field compared -- hubMatrix, tyreNdSlip and
    localAngularVelocity included."""
    _select_kernel(monkeypatch, kernel)
    from projectd_core_b200.pyprojectd import CAR_STATE_DTYPE
    n = 32
    b = _batch(oracle, n)
    assert b.tick_kernel_instance() == kernel
    starts = drive_start_states(oracle, lay, "driftplayground", 64)
    pick = [i for i in range(64) if (i // 16) % 4 in (1, 3)][:n]          # the rolling starts (10-16 m/s)
    refs = [oracle.RefSim() for _ in range(n)]
    for r, i in zip(refs, pick):
        r.set_state(starts[i][0]); r.set_time(0.0)
    recs = [r.state() for r in refs]
    obs = b.obs_tensor()
    worst = 0.0
    for t in range(120):
        for k, r in enumerate(refs):
            r.set_controls(**drive_controls(t, pick[k], lay, recs[k]))
        b.restore(np.stack([r.state() for r in refs], axis=1)); b.set_time(refs[0].time())
        b.step(DT, 1); b.observe(); b.sync()
        o = obs.cpu().numpy()
        for k, r in enumerate(refs):
            r.step(); recs[k] = r.state()
            if t % 4 != k % 4:
                continue
            want = np.zeros(664, np.uint8); r.L.pdref_get_car_state(r.h, want.ctypes.data)
            want = np.frombuffer(want, dtype=CAR_STATE_DTYPE)[0]
            got = np.frombuffer(b.car_state_bytes(k), dtype=CAR_STATE_DTYPE)[0]
            for name in CAR_STATE_DTYPE.names:
                if name in ("carId", "simId"):
                    continue
                g, w = got[name], want[name]
                if name == "controls":
                    for cn in g.dtype.names:
                        if cn == "isShifterSupported":
                            continue
                        assert abs(float(g[cn]) - float(w[cn])) <= 1e-6, (name, cn, g[cn], w[cn])
                    continue
                if np.issubdtype(g.dtype, np.integer):
                    assert np.array_equal(g, w), (t, k, name, g, w)
                    continue
                g = np.asarray(g, np.float64); w = np.asarray(w, np.float64)
                if name in ("bodyMatrix", "hubMatrix"):
                    gm = g.reshape(-1, 16); wm = w.reshape(-1, 16)
                    err = max(np.abs(gm[:, :12] - wm[:, :12]).max(), np.abs(gm[:, 12:15] - wm[:, 12:15]).max() / 10.0)   # positions: 1e-3 m in world coordinates of ~100 m (fp32 ulp 8e-6)
                elif name in ("bodyPos", "tyreContacts"):
                    err = np.abs(g - w).max() / 10.0
                elif g.ndim and name not in ("probes", "lookAhead", "tyreLoad", "tyreAngularSpeed", "tyreSlipRatio", "tyreNdSlip"):
                    err = np.abs(g - w).max() / max(float(np.linalg.norm(w)), 1.0)
                else:
                    err = (np.abs(g - w) / np.maximum(np.abs(w), 1.0)).max()
                worst = max(worst, float(err))
                assert err <= 3e-4, (t, k, name, g, w)
            ref_obs = np.concatenate([want["localVelocity"], want["localAngularVelocity"], want["tyreNdSlip"], [want["bodyVsTrack"], want["velocityVsTrack"]],
                                      want["lookAhead"], want["probes"][:7]]).astype(np.float64)
            sc = np.maximum(np.abs(ref_obs), 1.0)
            sc[0:3] = max(np.linalg.norm(ref_obs[0:3]), 1.0); sc[3:6] = max(np.linalg.norm(ref_obs[3:6]), 1.0)
            assert (np.abs(o[k] - ref_obs) / sc).max() <= 3e-4, (t, k, o[k], ref_obs)
    assert worst > 0.0


def test_env_config_knobs_are_live(oracle):
    """PdEnvConfig (ProjectDEnv's class attributes, projectd_env.py:27-53) reaches the kernels: gas range, termination switches,
    penalties, the clutch / gear overrides of auto_clutch / auto_shift off."""
    import torch
    from projectd_core_b200.env import BatchedProjectDEnv
    n = 64
    act = torch.zeros((n, 2), device="cuda"); act[:, 1] = -1.0          # a1 = -1 -> gas = min_gas
    lay = oracle.Layout(); o_gas = lay.fields["car.ctlGas"][0]; o_cl = lay.fields["car.ctlClutch"][0]; o_rg = lay.fields["car.ctlRequestedGear"][0]
    env = BatchedProjectDEnv(oracle.BASE_PATH, num_envs=n, device=0, min_gas=0.25, max_gas=0.75, auto_clutch=False, auto_shift=False,
                             terminate_when_stuck=True, stuck_timeout=0.05, terminate_stuck_penalty=7.0, autoreset_mode=1)
    env.reset()
    env.step(act); torch.cuda.synchronize()
    st = env.batch.snapshot()
    assert np.allclose(st[o_gas].view(np.float32), 0.25) and np.allclose(st[o_cl].view(np.float32), 1.0) and (st[o_rg].view(np.int32) == 2).all()
    act[:, 1] = 1.0
    env.step(act); torch.cuda.synchronize()
    assert np.allclose(env.batch.snapshot()[o_gas].view(np.float32), 0.75)
    # stuck after 0.05 s without reaching a new track point: every env terminates with the configured penalty
    done_any = torch.zeros(n, dtype=torch.bool, device="cuda"); pen = None
    act[:, 1] = -1.0
    for t in range(40):
        obs, rew, term, trunc, _ = env.step(act)
        if pen is None and bool(term.any()):
            pen = float(rew[term].max())
        done_any |= term
    assert bool(done_any.all()) and pen is not None and pen <= -7.0 + 0.2
    env.close()
    # with the switch off nothing terminates for being stuck
    env = BatchedProjectDEnv(oracle.BASE_PATH, num_envs=n, device=0, terminate_when_stuck=False, stuck_timeout=0.05)
    env.reset()
    act[:, 1] = -1.0
    seen = False
    for t in range(60):
        obs, rew, term, trunc, _ = env.step(act)
        seen |= bool(term.any())
    assert not seen
    env.close()


def test_env_step_is_ordered_with_callers_stream(oracle):
    """BatchedProjectDEnv.step on torch's default stream: actions produced just before the call and results consumed right after
    it, with no synchronisation by the caller, must equal a fully synchronised run (stream ordering)."""
    import torch
    from projectd_core_b200.env import BatchedProjectDEnv
    n = 512
    outs = []
    for sync in (False, True):
        env = BatchedProjectDEnv(oracle.BASE_PATH, num_envs=n, device=0, seed=11, teleport_mode=2, autoreset_mode=1)
        env.reset()
        g = torch.Generator(device="cuda"); g.manual_seed(4)
        tot = torch.zeros(n, device="cuda"); nd = torch.zeros(n, device="cuda"); acc = torch.zeros((n, 24), device="cuda")
        big = torch.empty((2048, 2048), device="cuda")
        for t in range(200):
            big.normal_(generator=g)                                        # work in flight on the caller's stream
            a = torch.tanh((big[:n, :2] * 0.7)).contiguous()               # actions produced by that work, then dropped
            if sync:
                torch.cuda.synchronize()
            obs, rew, term, trunc, _ = env.step(a)
            del a
            if sync:
                torch.cuda.synchronize()
            tot += rew; nd += term; acc += obs
        torch.cuda.synchronize()
        outs.append((tot.cpu().numpy(), nd.cpu().numpy(), acc.cpu().numpy(), env.batch.snapshot()))
        env.close()
    for x, y in zip(outs[0], outs[1]):
        assert np.array_equal(x, y)


def test_tunes_raw_and_unsupported(oracle):
    """setCarTune clamps / scales through the spinner, setCarRawTune writes the raw value (SetupManager.cpp:276-288,394-399);
    per-wheel front suspension tunes are live; reference variables this build cannot honour fail loudly."""
    from projectd_core_b200 import Batch, PdError
    b = _batch(oracle, 2)
    r = oracle.RefSim()                 # same env-style set-up on both sides (the reference's scoring variables are process-global)
    before = b.params_bytes().copy()
    for name, val in (("FRONT_BIAS", 55.0), ("SPRING_RATE_LF", 30.0), ("DAMP_BUMP_RF", 2500.0), ("CAMBER_LF", -20.0), ("TOE_OUT_RF", 12.0),
                      ("ROD_LENGTH_LF", 50.0), ("ARB_FRONT", 12000.0), ("INTERNAL_GEAR_2", 3.2), ("WING_1", 3.0)):
        b.set_tune(name, val); r.L.pdref_set_tune(r.h, name.encode(), val)
    ref = r.params_bytes(); mine = b.params_bytes()[: len(ref)]
    assert np.array_equal(mine, ref), np.nonzero(mine != ref)[0][:20]
    assert (b.params_bytes() != before).sum() >= 20, "the tunes must really have changed the parameter block"
    b3 = Batch(oracle.BASE_PATH, n_envs=2, device=0); b3.set_raw_tune("FRONT_BIAS", 0.61)
    b4 = Batch(oracle.BASE_PATH, n_envs=2, device=0); b4.set_tune("FRONT_BIAS", 61.0)
    assert np.array_equal(b3.params_bytes(), b4.params_bytes())
    for name in ("SPRING_RATE_LR", "DAMP_REBOUND_RR", "TURBO_0", "CENTER_DIFF_POWER"):
        with pytest.raises(PdError):
            b.set_tune(name, 1.0)
    b.set_tune("NO_SUCH_VARIABLE", 1.0)        # unknown to the reference as well: ignored there, ignored here


def _wall_approach_states(oracle, lay, n):
    """n oracle states a few ticks before a car first touches a wall (different start points / steering), found by driving."""
    key = ("wall", n)
    if key in _WALL_CACHE:
        return _WALL_CACHE[key]
    out = []
    k = 0
    while len(out) < n and k < 4 * n:
        r = oracle.RefSim(); r.set_collision_response(True)
        r.teleport_spline(0.05 + 0.9 * ((k * 0.37) % 1.0))
        steer = (0.25 if k % 2 else -0.25) * (1 + 0.2 * (k % 3)); gas = 0.7 + 0.1 * (k % 3)
        hist = []
        for t in range(2200):
            r.set_controls(steer=steer, gas=gas)
            hist.append((r.state().copy(), r.time()))
            r.step()
            if lay.get(r.state(), "car.collisionFlag") and len(r.contacts()) > 0:
                rec, tm = hist[max(0, len(hist) - 12)]
                out.append((rec, tm, steer, gas))
                break
        r.close(); k += 1
    _WALL_CACHE[key] = out
    return out


_WALL_CACHE = {}


@pytest.mark.parametrize("kernel", ["k_tick_quad<4>", "k_tick"])
def test_collision_response_tracks_oracle(oracle, lay, kernel, monkeypatch):
    """SURVEY.md A14 response: cars that run into walls with auto-teleport off.  (a) single tick from identical states on the odd
    frames where the contact joints are made: same contact set (count, kind; positions / normals / depths to 1e-4) and the
    state within the single-tick rule; (b) one second free running through the impact and the slide along the wall: the GPU car
    stays with the oracle's (bound 15 cm / 3 deg after the impact; the oracle's car bounces off instead of driving through)."""
    _select_kernel(monkeypatch, kernel)
    n = 16
    starts = _wall_approach_states(oracle, lay, n)
    assert len(starts) == n
    b = _batch(oracle, n)
    assert b.tick_kernel_instance() == kernel
    refs = [oracle.RefSim() for _ in range(n)]
    for r, (rec, tm, steer, gas) in zip(refs, starts):
        r.set_collision_response(True); r.set_state(rec); r.set_time(0.0); r.set_controls(steer=steer, gas=gas)
    # (a) identical states, 60 ticks around the first impact
    ncont = 0; nodd = 0
    for t in range(60):
        before = [r.state() for r in refs]
        tb = refs[0].time()
        b.restore(np.stack(before, axis=1)); b.set_time(tb)
        b.step(DT, 1)
        out = b.snapshot()
        for i, r in enumerate(refs):
            odd = lay.get(before[i], "car.physFrame") & 1
            live = r.contacts() if not odd else None          # joints that stay alive for this (even) frame in the free-running oracle
            r.step()
            ref = r.state()
            if odd:
                nodd += 1
                want = r.contacts(); got = b.contacts(i)
                assert len(got) == min(len(want), 8), (t, i, got, want)
                if len(want):
                    ncont += 1
                    assert np.array_equal(got[:, 7], want[:len(got), 7])
                    assert np.abs(got[:, :7] - want[:len(got), :7]).max() <= 1e-4 * max(1.0, float(np.abs(want[:, :3]).max())), (t, i, got, want)
                bad, w = compare_records(lay, out[:, i], ref, tol=1e-4)
                if bad:
                    left = arbitrate(oracle, lay, "driftplayground", before[i], tb, ref, bad)
                    # the arbiter's oracle run starts without live joints too (a restored state has none), as the GPU's does
                    assert not left, (t, i, left[:4])
    assert ncont >= n, "the approach states must lead into contacts (%d contact frames of %d odd frames)" % (ncont, nodd)
    # (b) one second free running from the approach states
    for r, (rec, tm, steer, gas) in zip(refs, starts):
        r.set_state(rec); r.set_time(0.0); r.set_controls(steer=steer, gas=gas)
    b.restore(np.stack([s[0] for s in starts], axis=1)); b.set_time(0.0)
    ctl = np.zeros((n, 5), np.float32)
    for i, s in enumerate(starts):
        ctl[i, 0] = s[2]; ctl[i, 4] = s[3]
    hit_ticks = 0
    for t in range(333):
        for r in refs:
            r.step()
        b.set_controls(ctl, None, smooth=True)
        b.step(DT, 1)
        hit_ticks += sum(1 for r in refs if len(r.contacts()) > 0)
    out = b.snapshot()
    assert hit_ticks >= 20 * n, "the cars must really spend time in contact (%d car-ticks)" % hit_ticks
    worst_p = worst_a = 0.0
    for i, r in enumerate(refs):
        ref = r.state()
        dp = [lay.get(out[:, i], "chassis." + k) - lay.get(ref, "chassis." + k) for k in ("px", "py", "pz")]
        dq = [lay.get(out[:, i], "chassis." + k) - lay.get(ref, "chassis." + k) for k in ("qw", "qx", "qy", "qz")]
        worst_p = max(worst_p, math.sqrt(sum(d * d for d in dp))); worst_a = max(worst_a, 2 * math.sqrt(sum(d * d for d in dq)))
    print("collision response, 1 s free run: worst position %.4f m, heading %.3f deg" % (worst_p, math.degrees(worst_a)))
    # measured on B200: 9.5 cm / 2.3 deg worst over 16 impacts (a wall impact amplifies round-off like any stiff event)
    assert worst_p <= 0.15 and worst_a <= math.radians(3.0), (worst_p, math.degrees(worst_a))


@pytest.mark.parametrize("track", ["driftplayground", "yamanashi_short", "ebisu_touge", "euphoria_hillside_park"])
def test_compute_fat_points_matches_shipped_spline_cache(oracle, track):
    """Track::computeFatPoints + computeSideLocation (Sim/Track.cpp:366-467) as batch ray casting on the GPU,
    against the reference-held spline.cache of every track that ships surfaces.bin -- a known-answer test for the general
    (slanted) rays of the side traces (up to 1000 dependent rays per point and side), not just the vertical ones."""
    from projectd_core_b200.binding import compute_fat_points
    base = oracle.BASE_PATH + "/content/tracks/%s/" % track
    want = np.fromfile(base + "spline.cache", dtype=np.float32).reshape(-1, 15)
    got = compute_fat_points(oracle.BASE_PATH, track)
    assert got.shape == want.shape
    err_best = np.abs(got[:, 0:3] - want[:, 0:3]).max(axis=1)
    err_side = np.maximum(np.abs(got[:, 3:6] - want[:, 3:6]).max(axis=1), np.abs(got[:, 6:9] - want[:, 6:9]).max(axis=1))
    err_dir = np.abs(got[:, 12:15] - want[:, 12:15]).max(axis=1)
    print("%s: %d points; best max %.2e; sides: max %.3f m, %d of %d points beyond 2 cm; forwardDir max %.2e" % (track, len(want), err_best.max(), err_side.max(), int((err_side > 0.02).sum()), len(want), err_dir.max()))
    # forwardDir: identical except at the LAST point, where the shipped caches hold the wrap-around direction of an older
    # computeFatPoints (the commented-out variant at Track.cpp:397-400): a few 1e-3 there
    assert err_best.max() <= 1e-4 and err_dir[:-1].max() <= 1e-5 and err_dir[-1] <= 2e-2
    # a side trace stops at the first ray whose hit breaks the continuity rules: one trace step (1 cm) of slack, and a few points
    # where ODE's any-hit-per-mesh ray and the closest-hit ray used here see different triangles at an overlap of meshes
    assert (err_side > 0.02).mean() <= 0.02, (int((err_side > 0.02).sum()), len(want))
    assert np.abs(got[:, 9:12] - 0.5 * (got[:, 3:6] + got[:, 6:9])).max() <= 1e-5


def test_track_without_spline_cache_loads(oracle, tmp_path):
    """A track whose spline.cache is missing loads anyway: pd_create regenerates the fat points on the GPU."""
    import shutil
    from projectd_core_b200 import Batch
    base = tmp_path / "base"
    shutil.copytree(oracle.BASE_PATH, base)
    os.remove(base / "content" / "tracks" / "driftplayground" / "spline.cache")
    b = make_env_like(Batch(str(base), n_envs=8, device=0))
    ref = _batch(oracle, 8)
    assert b.track_info()["nFatPoints"] == ref.track_info()["nFatPoints"]
    b.teleport_spline(np.linspace(0, 0.9, 8)); ref.teleport_spline(np.linspace(0, 0.9, 8))
    b.step(DT, 50); ref.step(DT, 50)
    sa, sb = b.snapshot(), ref.snapshot()
    lay = oracle.Layout()
    for i in range(8):
        bad, w = compare_records(lay, sa[:, i], sb[:, i], tol=1e-3)
        assert not [x for x in bad if math.isinf(x[3])], bad[:5]


def test_vector_env_interface(oracle):
    """the gymnasium VectorEnv face of the batched env (next-step auto-reset): spaces, reset / step signatures,
    terminal observation on the finishing step, reset observation one step later, truncation bookkeeping."""
    import torch
    from projectd_core_b200.vector_env import make_vec
    n = 128
    env = make_vec(oracle.BASE_PATH, num_envs=n, device=0, seed=3, teleport_mode=2)
    assert env.num_envs == n and env.single_observation_space.shape == (24,) and env.single_action_space.shape == (2,)
    assert env.observation_space.shape == (n, 24) and env.action_space.shape == (n, 2)
    obs, info = env.reset(seed=5)
    assert obs.shape == (n, 24) and obs.is_cuda and info == {}
    seen_done = torch.zeros(n, dtype=torch.bool, device="cuda"); after_reset_ok = 0
    prev_done = torch.zeros(n, dtype=torch.bool, device="cuda")
    for t in range(1200):
        a = torch.rand((n, 2), device="cuda") * 2 - 1
        obs, rew, term, trunc, info = env.step(a)
        assert obs.shape == (n, 24) and rew.shape == (n,) and term.dtype == torch.bool and trunc.dtype == torch.bool
        # the step after a terminal one is the reset step: reward 0, not done, car (nearly) at rest on the track
        if bool(prev_done.any()):
            assert float(rew[prev_done].abs().max()) == 0.0 and not bool(term[prev_done].any())
            assert float(obs[prev_done, 0:3].abs().max()) < 1.0
            after_reset_ok += int(prev_done.sum())
        prev_done = term | trunc
        seen_done |= term
    assert int(seen_done.sum()) >= n // 16 and after_reset_ok > 0
    env.close()


@pytest.mark.parametrize("kernel", ["k_tick_quad<4>", "k_tick"])
def test_brake_disc_temperatures_and_ebb(oracle, lay, kernel, monkeypatch, tmp_path, hostsim):
    """BrakeSystem's optional parts on the GPU (disc temperatures BrakeSystem.cpp:151-168, EBBMode::Internal :95-118), on a derived
    car whose brakes.ini carries [TEMPS_*] / [EBB] (no bundled car ships that data): single-tick rule on both tick kernels."""
    from projectd_core_b200 import Batch
    from parity_util import make_brake_temps_base, params_equal
    _select_kernel(monkeypatch, kernel)
    base, car = make_brake_temps_base(tmp_path, oracle.BASE_PATH)
    n = 32
    b = make_env_like(Batch(base, n_envs=n, device=0, car=car))
    refs = [oracle.RefSim(car=car, base=base) for _ in range(n)]
    assert params_equal(b.params_bytes(), refs[0].params_bytes(), hostsim)
    for i, r in enumerate(refs):
        r.teleport_spline(i / n)
    hot = 0.0
    for t in range(500):
        for i, r in enumerate(refs):
            brake = 0.3 + 0.02 * i if (t % 250) > 170 else 0.0
            r.set_controls(steer=0.2 * math.sin(0.01 * t + i), gas=0.0 if brake else 1.0, brake=brake)
        before = [r.state() for r in refs]; tb = refs[0].time()
        b.restore(np.stack(before, axis=1)); b.set_time(tb); b.step(DT, 1)
        out = b.snapshot()
        for i, r in enumerate(refs):
            r.set_state(before[i]); r.step()
            ref = r.state()
            bad, w = compare_records(lay, out[:, i], ref, tol=1e-4)
            if bad:
                left = arbitrate(oracle, lay, "driftplayground", before[i], tb, ref, bad, car=car, base=base)
                assert not left, (t, i, left[:5])
            hot = max(hot, lay.get(out[:, i], "car.brakeDiscT0"))
    assert hot > 20.01


@pytest.mark.parametrize("kind,kernel", [("ml", "k_tick_quad<4>"), ("ml", "k_tick"), ("heave", "k_tick_quad<8>"), ("heave", "k_tick"), ("throttle", "k_tick_quad<4>")])
def test_multilink_and_heave_spring_variants(oracle, lay, kind, kernel, monkeypatch, tmp_path, hostsim):
    """SuspensionML (SuspensionML.cpp:15-137) and HeaveSpring (HeaveSpring.cpp:11-149) on the GPU, on cars derived from ks_mazda_rx7_tuned (no bundled
    car ships such data; parity_util.make_variant_car_base): parameter block = reference init, single-tick rule on a 4-lanes-per-car instance and on
    the thread-per-car instance."""
    from projectd_core_b200 import Batch
    from parity_util import make_variant_car_base, params_equal
    _select_kernel(monkeypatch, kernel)
    base, car = make_variant_car_base(tmp_path, oracle.BASE_PATH, kind)
    n = 32
    b = make_env_like(Batch(base, n_envs=n, device=0, car=car))
    assert b.tick_kernel_instance() == ("k_tick<dwb,dwb>" if kernel == "k_tick" else kernel[:-1] + ",dwb,dwb>")
    refs = [oracle.RefSim(car=car, base=base) for _ in range(n)]
    assert params_equal(b.params_bytes(), refs[0].params_bytes(), hostsim)
    for i, r in enumerate(refs):
        r.teleport_spline(i / n)
    for t in range(260):
        for i, r in enumerate(refs):
            r.set_controls(**drive_controls(t, i + 16 * (i % 3), lay, r.state()))
        before = [r.state() for r in refs]; tb = refs[0].time()
        b.restore(np.stack(before, axis=1)); b.set_time(tb); b.step(DT, 1)
        out = b.snapshot()
        for i, r in enumerate(refs):
            r.set_state(before[i]); r.step()
            ref = r.state()
            bad, w = compare_records(lay, out[:, i], ref, tol=1e-4)
            if bad:
                left = arbitrate(oracle, lay, "driftplayground", before[i], tb, ref, bad, car=car, base=base)
                assert not left, (t, i, left[:5])


def test_device_bvh_rays_equal_host_bvh_rays(oracle, monkeypatch):
    """the ray caster's tree built on the device (linear BVH: Morton keys, radix sort, Karras' internal nodes, bottom-up refit;
    csrc/pd_lbvh.h).  Closest-hit rays must not depend on the tree they walk: 60 000 general rays (slanted, vertical, long, missing) give the
    same hit flags, surfaces and distances on the device-built and on the host-built tree, and the reference-held spline.cache is regenerated
    bit for bit through the device-built one (Track::computeFatPoints, the KAT of test_compute_fat_points_matches_shipped_spline_cache)."""
    from projectd_core_b200 import Batch
    rng = np.random.default_rng(17)
    fat = np.fromfile(oracle.BASE_PATH + "/content/tracks/driftplayground/spline.cache", dtype=np.float32).reshape(-1, 15)
    n = 60000
    idx = rng.integers(0, len(fat), n)
    rays = np.zeros((n, 7), np.float32)
    rays[:, 0:3] = fat[idx, 0:3] + rng.uniform(-15, 15, (n, 3)).astype(np.float32) * np.array([1, 0.2, 1], np.float32) + np.array([0, 3.0, 0], np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32); d[:, 1] = -np.abs(d[:, 1]) - 0.2; d[: n // 4] = np.array([0, -1, 0], np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays[:, 3:6] = d; rays[:, 6] = rng.uniform(0.5, 60.0, n).astype(np.float32)
    monkeypatch.setenv("PD_DEVICE_BVH", "0")
    bh = make_env_like(Batch(oracle.BASE_PATH, n_envs=1, device=0))
    assert bh.bvh_info()[0] is False
    host = bh.raycast(rays)
    monkeypatch.setenv("PD_DEVICE_BVH", "1")
    bd = make_env_like(Batch(oracle.BASE_PATH, n_envs=1, device=0))
    on_dev, nodes, depth = bd.bvh_info()
    assert on_dev and nodes == 2 * bd.track_info()["nTris"] - 1 and 10 < depth <= 45, (on_dev, nodes, depth)
    dev = bd.raycast(rays)
    assert np.array_equal(dev[:, 0], host[:, 0]) and 0.2 < host[:, 0].mean() < 0.98, host[:, 0].mean()
    hit = host[:, 0] == 1
    assert np.array_equal(dev[hit, 7], host[hit, 7]), "surface ids differ"
    assert np.array_equal(dev[hit, 1:7], host[hit, 1:7]), "hit positions / normals differ"
    # the reference-held known answers through the device-built tree
    from projectd_core_b200.binding import compute_fat_points
    mine = compute_fat_points(oracle.BASE_PATH, "driftplayground", device=0)
    assert np.array_equal(mine[:, 0:3], fat[:, 0:3])


def test_full_size_batch_65536_envs(oracle, lay, monkeypatch):
    """BASELINE configs[2]'s per-GPU load (65536 envs, the bench configuration: random actions, next-step auto-reset, an env
    terminates on a hit) through size-independent properties: (1) the rollout stays finite and the episode counters add up;
    (2) sharding: the first and the last 4096 envs of the batch equal, bit for bit, 4096-env batches seeded at the same global
    env ids (env-id keyed RNG, the same kernel instance); (3) sampled oracle parity at full size: after the rollout, one more
    tick of 64 envs picked across the batch against the oracle started from those envs' records (single-tick rule)."""
    import torch
    from projectd_core_b200 import Batch
    from projectd_core_b200.env import configure_like_env
    N, S, T = 65536, 4096, 500
    monkeypatch.setenv("PD_QUAD_MAX_ENVS", "0"); monkeypatch.setenv("PD_SERIAL_WIDE", "0")     # the shards run what the full batch runs

    def make(n, off):
        b = configure_like_env(Batch(oracle.BASE_PATH, n_envs=n, device=0))
        b.set_seed(1234, off); b.teleport_mode(2); b.set_autoreset(1)
        return b
    full = make(N, 0); lo = make(S, 0); hi = make(S, N - S)
    assert full.tick_kernel_instance() == "k_tick" and lo.tick_kernel_instance() == "k_tick"
    gen = torch.Generator(device="cuda"); gen.manual_seed(7)
    rew = torch.zeros(N, device="cuda"); done = torch.zeros(N, device="cuda", dtype=torch.int32)
    rs = torch.zeros(S, device="cuda"); ds = torch.zeros(S, device="cuda", dtype=torch.int32)
    ndone = torch.zeros((), device="cuda", dtype=torch.int64)
    for t in range(T):
        if t % 33 == 0:
            act = (torch.rand((N, 2), device="cuda", generator=gen) * 2 - 1).contiguous()
            a_lo = act[:S].contiguous(); a_hi = act[N - S:].contiguous()
        full.env_step(act, DT, None, rew, done); lo.env_step(a_lo, DT, None, rs, ds); hi.env_step(a_hi, DT, None, rs, ds)
        full.sync(); lo.sync(); hi.sync()
        ndone += (done != 0).sum()
    st = full.env_stats(reset=False)
    assert st[7] == 0, "non-finite cars"                      # nan counter
    assert int(st[0]) == int(ndone.item()) and st[0] == st[3] + st[4] + st[5] + st[6] + st[7], st
    assert st[0] > N * 0.02, "the rollout never reached episode ends"
    before = full.snapshot()
    assert np.isfinite(before[lay.fields["chassis.px"][0]].view(np.float32)).all()
    assert np.array_equal(before[:, :S], lo.snapshot()) and np.array_equal(before[:, N - S:], hi.snapshot())
    # sampled oracle parity: restore (drops nothing but live contact joints, which the bench configuration never has), one plain tick
    tb = full.time()
    full.restore(before); full.set_time(tb); full.step(DT, 1)
    after = full.snapshot()
    rng = np.random.default_rng(3)
    pick = sorted(set([0, 63, 64, N - 1] + rng.integers(0, N, 60).tolist()))
    cf = lay.fields["car.collisionFlag"][0]
    worst = 0.0; hits = 0; failures = []
    r = oracle.RefSim()
    for i in pick:
        r.set_state(before[:, i]); r.set_time(tb); r.step()
        ref = r.state()
        assert int(after[cf, i]) == int(ref[cf]), ("collision flag", i)
        if int(ref[cf]):            # the oracle answers a hit with contact joints, the env configuration with termination
            hits += 1; continue
        bad, w = compare_records(lay, after[:, i], ref, tol=1e-4)
        if bad:
            left = arbitrate(oracle, lay, "driftplayground", before[:, i], tb, ref, bad)
            if left:
                failures.append((i, left[:4]))
        else:
            worst = max(worst, w)
    assert not failures, failures
    assert hits < len(pick) // 2
