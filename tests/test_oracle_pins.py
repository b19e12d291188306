"""Independent pins of the ODE 0.16.3 restatement (oracle/ode_restate): quantities that hold for ANY correct implementation of
the rigid-body / joint step, whatever its internals -- SURVEY.md 8(c)'s list.  The restatement itself cannot be pinned against
ODE (source and binary are absent from the reference tree: DESIGN.md 3); these tests bound how wrong it could be."""
import ctypes
import math

import numpy as np
import pytest

from parity_util import centering_steer


def _fn(oracle, name, restype, argtypes):
    f = getattr(oracle.lib(), name); f.restype = restype; f.argtypes = argtypes
    return f


def test_constraints_hold_after_every_step(oracle):
    """||C(q)|| after the step: ball / fixed / slider anchors stay together and dball links keep their length while the car is
    launched, steered and braked for 6 s."""
    errs = _fn(oracle, "pdref_constraint_errors", None, [ctypes.c_void_p, ctypes.c_void_p])
    lay = oracle.Layout()
    r = oracle.RefSim(); r.teleport_spline(0.2)
    out = np.zeros(3, np.float64); worst_lin = worst_db = 0.0
    for t in range(2000):
        rec = r.state()
        steer = centering_steer(lay, rec) + 0.4 * math.sin(t / 37.0)
        r.set_controls(steer=max(-1.0, min(1.0, steer)), gas=1.0 if (t // 300) % 2 == 0 else 0.0, brake=0.0 if (t // 300) % 2 == 0 else 0.7)
        r.step()
        errs(r.h, out.ctypes.data)
        if t >= 30:      # the reference's teleport seats hubs / struts a few cm off their anchors (SuspensionStrut::setPositions); ERP closes that in ~10 steps
            worst_lin = max(worst_lin, out[0]); worst_db = max(worst_db, out[1])
    assert out[2] == 16                      # fixed + 2 x (3 dball + slider + ball) + 5 dball = 16 joints, 33 rows
    assert lay.get(r.state(), "car.speed") > 1.0
    assert worst_lin < 1e-3 and worst_db < 1e-2, (worst_lin, worst_db)     # measured: 0.3 mm on the ball / fixed / slider anchors, 5 mm on the loaded dball links (ERP 0.3 at speed)


def test_free_body_conserves_momentum():
    """dxStepIsland / dxStepBody on a free tumbling chassis (no joints, no gravity): linear momentum exact, angular momentum and
    energy drift small over 3 s (ODE's implicit gyroscopic term damps the energy slightly, never grows it wildly)."""
    import pdref
    f = _fn(pdref, "pdref_free_body_drift", None, [ctypes.c_int, ctypes.c_double, ctypes.c_void_p])
    out = np.zeros(3, np.float64)
    f(999, 1.0 / 333.0, out.ctypes.data)
    assert out[2] < 1e-5, out                   # linear velocity unchanged
    assert out[0] < 2e-2 and out[1] < 2e-2, out  # |L| and E within 2 % after 999 steps of a 1.4 rad/s tumble


def test_weight_is_carried_by_the_tyres(oracle):
    """Vertical force balance through the whole joint tree: cruising at ~7 m/s, the vertical tyre force averaged over 4.5 s equals
    the weight of the seven bodies (chassis 837.4 kg + fuel tank 22.2 + 2 hubs 28.8 + 2 strut bodies 7.2 + axle 100 = 1031.6 kg:
    car.ini TOTALMASS 995 = chassis + hubs + axle, Car.cpp:589-620; SuspensionStrut.cpp:132-137) within 4 %: gravity enters at the
    bodies, the only way to the ground is joints -> hubs / axle -> tyres.  (At a standstill the reference's sleep logic --
    body->stop() every tick after 50 quiet frames, Car.cpp:536-540 -- freezes the chassis wherever it was, so rest is no
    equilibrium to test against.)"""
    lay = oracle.Layout()
    r = oracle.RefSim(); r.teleport_spline(0.0)
    acc = []
    for t in range(3000):
        rec = r.state()
        r.set_controls(steer=centering_steer(lay, rec), gas=0.35 if lay.get(rec, "car.speed") < 5 else 0.12); r.step()
        if t >= 1500:
            rec = r.state()
            acc.append(sum(lay.get(rec, "tyre%d.load" % w) * lay.get(rec, "tyre%d.normalY" % w) for w in range(4)))
    weight = 9.80665 * (837.4 + 22.2 + 2 * 28.8 + 2 * 7.2 + 100.0)
    assert lay.get(r.state(), "car.speed") > 3.0 and not lay.get(r.state(), "car.collisionFlag")
    assert abs(np.mean(acc) - weight) / weight < 0.04, (np.mean(acc), weight)


def test_ldlt_solution_against_fp64_dense_resolve(oracle):
    """The restated single-precision LDL^T against a double-precision Gaussian elimination of the SAME 33 x 33 system, on states
    from a drive: the multipliers agree to 1e-3 of the largest one (the system's conditioning at fp32), median far below."""
    keep = _fn(oracle, "pdref_keep_system", None, [ctypes.c_void_p, ctypes.c_int])
    res = _fn(oracle, "pdref_resolve_fp64", None, [ctypes.c_void_p, ctypes.c_void_p])
    lay = oracle.Layout()
    r = oracle.RefSim(); r.teleport_spline(0.4); keep(r.h, 1)
    out = np.zeros(2, np.float64); rel = []
    for t in range(900):
        rec = r.state()
        r.set_controls(steer=centering_steer(lay, rec), gas=0.9); r.step()
        if t % 10 == 0:
            res(r.h, out.ctypes.data); rel.append(out[0])
            assert out[1] == 33
    rel = np.array(rel)
    assert rel.min() >= 0 and np.median(rel) < 1e-4 and rel.max() < 1e-3, (np.median(rel), rel.max())
