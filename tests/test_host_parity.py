"""CPU-side checks of the product's arithmetic and host logic (no GPU needed).

tests/hostsim compiles the SAME device functions the CUDA kernels are made of (projectd_core_b200/csrc/*.h) with
g++; it is a debugging aid, not a product path.  Here it is checked against the committed golden vectors of the
oracle (tests/golden) so that a kernel regression is caught before spending GPU time; the GPU tests proper
(tests/test_gpu_parity.py) run the CUDA library through the C ABI."""
import ctypes
import math

import numpy as np
import pytest

import pdref
from parity_util import compare_records

DT = 1.0 / 333.0


@pytest.fixture(scope="module")
def lay():
    return pdref.Layout()


def test_loader_params_equal_reference_init(hostsim, hostsim_env, golden):
    """car .ini/.lut/.rto parsing + setup spinners + tune re-seat == the reference's own init, byte for byte."""
    n = hostsim.hs_params_bytes()
    assert n == golden["params"].size
    buf = np.zeros(n, np.uint8)
    hostsim.hs_get_params(hostsim_env, buf.ctypes.data)
    diff = np.nonzero(buf != golden["params"])[0]
    assert diff.size == 0, ("parameter block differs at bytes", diff[:16])


def test_track_loader_matches_reference(hostsim, hostsim_env, golden):
    ti = np.zeros(16, np.uint32)
    hostsim.hs_get_track_info(hostsim_env, ti.ctypes.data)
    ref = golden["track_info"]
    assert ti[1] == ref[1] and ti[3] == ref[3] and ti[4] == ref[4]      # nTris, nFatPoints, nSplineNodes
    assert ti[8] == ref[8]                                              # computedTrackLength (float bits)


@pytest.mark.parametrize("variant", ["serial", "quad"])
def test_single_tick_vs_golden_pairs(hostsim, hostsim_env, golden, lay, variant):
    """One tick from the oracle's state (controls included) with the kernel arithmetic: ints exact, floats within
    1e-4 (north-star single-tick tolerance); rare exceedances up to 3e-4 are solver conditioning."""
    tick = hostsim.hs_tick if variant == "serial" else hostsim.hs_tick_quad
    worst = 0.0; nbad = 0
    step = 1 if variant == "serial" else 2
    idx = range(0, len(golden["pair_before"]), step)
    for k in idx:
        rec = golden["pair_before"][k].copy()
        tick(hostsim_env, rec.ctypes.data, DT, float(golden["pair_time"][k]))
        bad, w = compare_records(lay, rec, golden["pair_after"][k], tol=1e-4)
        assert not any(math.isinf(b[3]) for b in bad), (k, [b for b in bad if math.isinf(b[3])][:5])
        worst = max(worst, w); nbad += bool(bad)
    assert worst <= 3e-4, worst
    assert nbad <= max(1, len(idx) // 100), nbad


def test_quad_equals_serial_bitwise_mostly(hostsim, hostsim_env, golden, lay):
    """The 4-lane distribution changes only summation order of the chassis force / Schur terms: <= 2e-6 apart."""
    for k in range(0, len(golden["pair_before"]), 16):
        a = golden["pair_before"][k].copy(); b = a.copy()
        hostsim.hs_tick(hostsim_env, a.ctypes.data, DT, float(golden["pair_time"][k]))
        hostsim.hs_tick_quad(hostsim_env, b.ctypes.data, DT, float(golden["pair_time"][k]))
        bad, w = compare_records(lay, b, a, tol=2e-5)
        assert not bad, (k, bad[:5])


def test_teleport_vs_golden(hostsim, hostsim_env, golden, lay):
    """Car::teleportToSpline + reset (Car.cpp:1325-1340, 385-410) on the device functions."""
    for u, ref in zip(golden["tele_u"], golden["tele_state"]):
        rec = golden["tele_state"][0].copy()
        pid = hostsim.hs_point_id_at_distance(hostsim_env, float(u))
        hostsim.hs_teleport_point(hostsim_env, rec.ctypes.data, pid, 0.0)
        bad, w = compare_records(lay, rec, ref, tol=1e-6)
        assert not bad, (u, bad[:5])


def test_raycast_vs_golden(hostsim, hostsim_env, golden):
    rays = np.ascontiguousarray(golden["rays"]); ref = golden["ray_hits"]
    out = np.zeros_like(ref)
    hostsim.hs_raycast(hostsim_env, len(rays), rays.ctypes.data, out.ctypes.data)
    assert np.array_equal(out[:, 0], ref[:, 0]), "hit flags must match exactly"
    hit = ref[:, 0] == 1
    assert hit.sum() > 1000
    assert np.array_equal(out[hit, 7], ref[hit, 7]), "surface ids must match exactly"
    assert np.abs(out[hit, 1:7] - ref[hit, 1:7]).max() <= 2e-6


def test_sctm_vs_golden(hostsim, hostsim_env, golden):
    """SCTM::solve (Tyre/SCTM.cpp) restated in pd_car.h: same inputs -> outputs within 2e-6 relative."""
    x = np.ascontiguousarray(golden["sctm_in"])
    for w in range(4):
        y = np.zeros((len(x), 7), np.float32)
        hostsim.hs_sctm_solve(hostsim_env, w, len(x), x.ctypes.data, y.ctypes.data)
        ref = golden["sctm_out"][w]
        err = np.abs(y - ref) / np.maximum(np.abs(ref), 1.0)
        assert err.max() <= 2e-6, (w, err.max())


def test_probe_grid_walk_equals_exhaustive_scan(hostsim, hostsim_env, content_base):
    """The grid-indexed probe walk / nearest point (pd_track.h) must give the reference's exhaustive-loop results
    (Track::rayCastTrackBounds Track.cpp:497-562, getPointIdAtLocation :579-596) at arbitrary poses."""
    fat = np.fromfile(content_base + "/content/tracks/driftplayground/spline.cache", dtype=np.float32).reshape(-1, 15)
    rng = np.random.default_rng(5)
    n = 20000
    idx = rng.integers(0, len(fat), n)
    poses = np.zeros((n, 4), np.float32)
    poses[:, 0:3] = fat[idx, 0:3] + rng.uniform(-16, 16, (n, 3)).astype(np.float32) * np.array([1, 0.05, 1], np.float32)
    poses[:, 3] = rng.uniform(-math.pi, math.pi, n)
    walk = np.zeros((n, 8), np.float32); brute = np.zeros((n, 8), np.float32)
    hostsim.hs_probe_compare(hostsim_env, n, poses.ctypes.data, walk.ctypes.data, brute.ctypes.data)
    ok = walk[:, :7] >= 0     # -1: the walk declined (left the indexed area) and the kernel falls back to the scan
    assert ok.mean() > 0.99
    assert np.array_equal(walk[:, :7][ok], brute[:, :7][ok])
    okp = walk[:, 7] >= 0
    assert okp.mean() > 0.99, "the ring search must cover cars up to 16 m off the line"
    assert np.array_equal(walk[okp, 7], brute[okp, 7])


def test_free_running_1s_vs_golden_trajectory(hostsim, hostsim_env, golden, lay):
    """configs[0] drive, free running from the teleport state for 250 ticks with the kernel arithmetic vs the
    oracle's stored record: bounded divergence (5 cm / 0.5 deg, see test_gpu_parity)."""
    rec = golden["traj_state"][0].copy()
    t0 = float(golden["traj_time"][0])
    for t in range(250):
        lay.set(rec, "car.ctlSteer", 0.3 * math.sin(2 * math.pi * t / 999.0))
        lay.set(rec, "car.ctlGas", 0.1 + 0.9 * min(1.0, t / 333.0))
        hostsim.hs_tick(hostsim_env, rec.ctypes.data, DT, t0 + t * DT)
    ref = golden["traj_state"][1]
    dp = [lay.get(rec, "chassis." + k) - lay.get(ref, "chassis." + k) for k in ("px", "py", "pz")]
    assert math.sqrt(sum(d * d for d in dp)) <= 0.05, dp
    assert lay.get(rec, "car.currentGear") == lay.get(ref, "car.currentGear")


@pytest.mark.parametrize("variant", ["serial", "quad"])
def test_collision_flag_vs_golden(hostsim, hostsim_env, golden, lay, variant):
    """Car::collisionFlag after one tick on an odd physics frame, for driving states with the whole car
    translated towards walls / into the ground (floor box vs TRACK meshes with the body-local normal filter, hull mesh
    vs WALL meshes): the flag must equal the oracle's on every case."""
    tick = hostsim.hs_tick if variant == "serial" else hostsim.hs_tick_quad
    flags = golden["coll_flag"]
    assert 0.1 < flags.mean() < 0.9
    wrong = []
    for k in range(0, len(flags), 1 if variant == "serial" else 3):
        rec = golden["coll_before"][k].copy()
        tick(hostsim_env, rec.ctypes.data, DT, float(golden["coll_time"][k]))
        if lay.get(rec, "car.collisionFlag") != int(flags[k]):
            wrong.append(k)
    assert not wrong, wrong[:10]


def test_no_collision_check_on_even_frames(hostsim, hostsim_env, golden, lay):
    """PhysicsEngineODE::collisionStep tests the static meshes on odd frames only (PhysicsEngineODE.cpp:230-236)."""
    ks = np.nonzero(golden["coll_flag"])[0][:20]
    for k in ks:
        rec = golden["coll_before"][k].copy()
        lay.set(rec, "car.physFrame", 2)
        hostsim.hs_tick(hostsim_env, rec.ctypes.data, DT, float(golden["coll_time"][k]))
        assert lay.get(rec, "car.collisionFlag") == 0 and lay.get(rec, "car.physFrame") == 3


OTHER_CARS = ["ks_mazda_rx7_tuned", "ks_toyota_supra_mkiv_drift", "dthwsh_mazda_rx7_fc3s_sr20", "gravygarage_street_ae86_readie"]


@pytest.mark.parametrize("car", OTHER_CARS)
def test_other_cars_loader_and_single_tick(hostsim, oracle, content_base, car):
    """SURVEY.md N1 on the CPU tier: the loader's parameter block for the double-wishbone / turbocharged cars equals the
    reference's own init byte for byte, and the device functions of their kernel instances (compiled for the host) follow
    the oracle tick by tick from identical states (SuspensionDW.cpp:202-272, Turbo.cpp:11-40)."""
    import math
    from parity_util import compare_records
    r = oracle.RefSim(car=car); r.set_collision_response(False)
    h = hostsim.hs_create(content_base.encode(), b"driftplayground", car.encode())
    assert h, "loader rejected " + car
    hostsim.hs_set_assists(h, 1, 1, 1)
    for k, v in oracle.ENV_TUNES.items():
        hostsim.hs_set_tune(h, k.encode(), v)
    for k, v in oracle.ENV_SCORING.items():
        hostsim.hs_set_scoring_var(h, k.encode(), v)
    from parity_util import params_equal
    ref = r.params_bytes()
    mine = np.zeros(hostsim.hs_params_bytes(), np.uint8); hostsim.hs_get_params(h, mine.ctypes.data)
    assert params_equal(mine, ref, hostsim), np.nonzero(mine != ref)[0][:10]
    r.teleport_spline(0.3)
    lay = oracle.Layout()
    boost = 0.0
    for t in range(240):
        r.set_controls(steer=0.4 * math.sin(0.01 * t), gas=min(1.0, 0.2 + t / 150.0) if t < 180 else 0.0, brake=0.0 if t < 180 else 0.6)
        before = r.state().copy(); tb = r.time()
        rec = before.copy()
        hostsim.hs_tick(h, rec.ctypes.data, 1.0 / 333.0, tb)
        recq = None
        if t % 4 == 0:        # the 4-lanes-per-car form (one joint group per lane, the tank's beside the RR hub's on lane 3) every 4th tick
            recq = before.copy(); hostsim.hs_tick_quad(h, recq.ctypes.data, 1.0 / 333.0, tb)
        r.step()
        for mine in (rec, recq):
            if mine is None:
                continue
            bad, worst = compare_records(lay, mine, r.state(), tol=1e-4)
            bad = [x for x in bad if not (x[0].endswith(".wz") and x[3] < 5e-4)]       # hub spin about its own axis, held by short links: the GPU tier arbitrates these records through the oracle
            assert not bad, (t, mine is recq, bad[:6])
        boost = max(boost, lay.get(rec, "car.turboBoost"))
    assert boost > 0.01
    hostsim.hs_destroy(h)


def test_brake_disc_temperatures_and_ebb(hostsim, oracle, content_base, tmp_path):
    """BrakeSystem's optional parts (disc temperatures BrakeSystem.cpp:151-168, EBBMode::Internal :95-118) on a derived car whose
    brakes.ini carries [TEMPS_*] / [EBB]: loader block = reference init, and the braking ticks follow the oracle (disc temperatures
    rise, per-wheel brake torques follow the performance curves)."""
    from parity_util import compare_records, make_brake_temps_base, params_equal
    base, car = make_brake_temps_base(tmp_path, content_base)
    r = oracle.RefSim(car=car, base=base); r.set_collision_response(False)
    h = hostsim.hs_create(base.encode(), b"driftplayground", car.encode())
    assert h
    hostsim.hs_set_assists(h, 1, 1, 1)
    for k, v in oracle.ENV_TUNES.items():
        hostsim.hs_set_tune(h, k.encode(), v)
    for k, v in oracle.ENV_SCORING.items():
        hostsim.hs_set_scoring_var(h, k.encode(), v)
    mine = np.zeros(hostsim.hs_params_bytes(), np.uint8); hostsim.hs_get_params(h, mine.ctypes.data)
    assert params_equal(mine, r.params_bytes(), hostsim)
    lay = oracle.Layout()
    r.teleport_spline(0.1)
    hot = 0.0
    for t in range(700):
        brake = 0.8 if (t % 350) > 250 else 0.0
        r.set_controls(steer=0.1 * math.sin(0.01 * t), gas=0.0 if brake else 1.0, brake=brake)
        rec = r.state().copy(); tb = r.time()
        forms = (hostsim.hs_tick, hostsim.hs_tick_quad) if t % 7 == 0 else (hostsim.hs_tick,)      # thread-per-car form every tick, the 4-lane form every 7th
        outs = []
        for fn in forms:
            mine = rec.copy(); fn(h, mine.ctypes.data, DT, tb); outs.append(mine)
        first = outs[0]
        r.step()
        for mine in outs:
            bad, worst = compare_records(lay, mine, r.state(), tol=1e-4)
            bad = [x for x in bad if not (x[0].endswith(".wz") and x[3] < 5e-4)]
            assert not bad, (t, bad[:6])
        hot = max(hot, lay.get(first, "car.brakeDiscT0"))
    assert hot > 20.01, "the discs never warmed up"      # 2 s of simulated time: a few hundredths of a kelvin above the 20 C ambient
    hostsim.hs_destroy(h)


@pytest.mark.parametrize("kind", ["ml", "heave", "throttle"])
def test_multilink_and_heave_spring_variants(hostsim, oracle, content_base, tmp_path, kind):
    """SuspensionML (SuspensionML.cpp:15-137: a hub on five distance joints with the world's ERP / CFM, spring force of either sign) and
    HeaveSpring (HeaveSpring.cpp:11-149: third spring on the mean travel of an axle's two hubs), on cars derived from ks_mazda_rx7_tuned
    (no bundled car ships such data): loader block = reference init, then tick by tick against the oracle in the thread-per-car form and
    (every 4th tick) the 4-lanes-per-car form."""
    from parity_util import compare_records, make_variant_car_base, params_equal
    base, car = make_variant_car_base(tmp_path, content_base, kind)
    r = oracle.RefSim(car=car, base=base); r.set_collision_response(False)
    h = hostsim.hs_create(base.encode(), b"driftplayground", car.encode())
    assert h, "loader rejected the %s variant" % kind
    hostsim.hs_set_assists(h, 1, 1, 1)
    for k, v in oracle.ENV_TUNES.items():
        hostsim.hs_set_tune(h, k.encode(), v)
    for k, v in oracle.ENV_SCORING.items():
        hostsim.hs_set_scoring_var(h, k.encode(), v)
    mine = np.zeros(hostsim.hs_params_bytes(), np.uint8); hostsim.hs_get_params(h, mine.ctypes.data)
    ref = r.params_bytes()
    assert params_equal(mine, ref, hostsim), np.nonzero(mine != ref)[0][:12]
    lay = oracle.Layout()
    r.teleport_spline(0.3)
    for t in range(320):
        r.set_controls(steer=0.4 * math.sin(0.01 * t), gas=min(1.0, 0.2 + t / 150.0) if t < 240 else 0.0, brake=0.0 if t < 240 else 0.6)
        before = r.state().copy(); tb = r.time()
        outs = []
        for fn in ((hostsim.hs_tick, hostsim.hs_tick_quad) if t % 4 == 0 else (hostsim.hs_tick,)):
            rec = before.copy(); fn(h, rec.ctypes.data, DT, tb); outs.append(rec)
        r.step()
        for rec in outs:
            bad, worst = compare_records(lay, rec, r.state(), tol=1e-4)
            bad = [x for x in bad if not (x[0].endswith(".wz") and x[3] < 5e-4)]
            assert not bad, (t, bad[:6])
    hostsim.hs_destroy(h)
