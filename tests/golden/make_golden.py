"""Generates tests/golden/demo_golden.npz from the ORACLE (oracle/_ref/libpdref.so = the reference's own
Car/Sim/Core sources compiled in place + the restated ODE back-end, see oracle/README.md).

Run here (container with /root/reference):   python tests/golden/make_golden.py
The vectors travel with the repo, so the hostsim / GPU tests can check against them where the reference tree is
absent.  Content (all from the bundled demo car ks_toyota_ae86_drift on driftplayground, env-style setup of
pyprojectd/projectd_env.py:118-136):

  params, track_info            reference init: car parameter block (include/pd_params.h layout) + track summary
  traj_tick/time/state          BASELINE configs[0]: scripted throttle/steer, 10 000 ticks, record every 250 ticks
  pair_before/after/time        single-tick pairs (state incl. controls before Simulator::step -> state after)
  tele_u/state                  teleportCarToSpline(u) states
  rays/ray_hits                 world rays vs the track trimesh
  sctm_in/out                   SCTM::solve inputs -> outputs for the 4 tyres
  coll_before/coll_flag         collision detection: driving states with the whole car translated (towards
                                walls / into the ground), physics frame odd -> Car::collisionFlag after one Simulator::step
"""
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle"))
import pdref  # noqa: E402

DT = 1.0 / 333.0


def scripted(t):
    """gas = 0.1+0.9*min(1,t/333), steer = 0.3*sin(2 pi t/999)."""
    return 0.3 * math.sin(2 * math.pi * t / 999.0), 0.1 + 0.9 * min(1.0, t / 333.0)


def main():
    assert pdref.available(), "build the oracle first: make -C oracle && make -C oracle content"
    out = {}
    r = pdref.RefSim(); r.set_collision_response(False)     # fixtures of the CPU tier: detection only (the host build of the kernels has no contact generator; the response is tested on the GPU, live against the oracle)
    out["params"] = r.params_bytes()
    ti = np.zeros(12, np.uint32); r.L.pdref_get_track_info(r.h, ti.ctypes.data); out["track_info"] = ti

    # ---- configs[0] trajectory ----
    r.teleport_spline(0.0)
    ticks, times, states = [], [], []
    for t in range(10000):
        if t % 250 == 0:
            ticks.append(t); times.append(r.time()); states.append(r.state().copy())
        steer, gas = scripted(t)
        r.set_controls(steer=steer, gas=gas)
        r.step(DT)
    ticks.append(10000); times.append(r.time()); states.append(r.state().copy())
    out["traj_tick"] = np.array(ticks, np.int32); out["traj_time"] = np.array(times, np.float64); out["traj_state"] = np.stack(states)

    # ---- single-tick pairs from 6 different drives ----
    rng = np.random.default_rng(20261017)
    before, after, ptime = [], [], []
    for drive in range(6):
        s = pdref.RefSim(); s.set_collision_response(False); s.teleport_spline(drive / 6.0)
        keep = set(rng.choice(1500, 40, replace=False).tolist()) | {0, 1, 2}
        for t in range(1500):
            steer = 0.5 * math.sin(2 * math.pi * t / (400.0 + 100 * drive) + drive)
            gas = min(1.0, 0.2 + 0.8 * t / 300.0) * (0.5 + 0.1 * drive)
            brake = 0.7 if (drive % 2 == 1 and 700 < t < 780) else 0.0
            hand = 1.0 if (drive == 4 and 900 < t < 940) else 0.0
            s.set_controls(steer=steer, gas=0.0 if brake > 0 else gas, brake=brake, hand_brake=hand)
            if t in keep:
                before.append(s.state().copy()); ptime.append(s.time())
            s.step(DT)
            if t in keep:
                after.append(s.state().copy())
        s.close()
    out["pair_before"] = np.stack(before); out["pair_after"] = np.stack(after); out["pair_time"] = np.array(ptime, np.float64)

    # ---- teleports ----
    us = np.array([0.0, 0.11, 0.25, 0.37, 0.5, 0.63, 0.81, 0.97], np.float32)
    tele = []
    for u in us:
        s = pdref.RefSim(); s.teleport_spline(float(u)); tele.append(s.state().copy()); s.close()
    out["tele_u"] = us; out["tele_state"] = np.stack(tele)

    # ---- rays ----
    fat = np.fromfile(pdref.BASE_PATH + "/content/tracks/driftplayground/spline.cache", dtype=np.float32).reshape(-1, 15)
    n = 4000
    idx = rng.integers(0, len(fat), n)
    rays = np.zeros((n, 7), np.float32)
    rays[:, 0:3] = fat[idx, 0:3] + rng.uniform(-12, 12, (n, 3)).astype(np.float32) * np.array([1, 0, 1], np.float32) + np.array([0, 2.5, 0], np.float32)
    rays[:, 4] = -1; rays[:, 6] = 3.0
    rays[n // 2:, 6] = 1000.0; rays[n // 2:, 1] += 10
    # a quarter of the rays are oblique (general BVH path)
    ob = slice(0, n // 4)
    d = rng.normal(size=(n // 4, 3)).astype(np.float32); d[:, 1] = -np.abs(d[:, 1]) - 0.3
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays[ob, 3:6] = d; rays[ob, 6] = 30.0
    hits = np.zeros((n, 8), np.float32)
    r.L.pdref_raycast(r.h, n, rays.ctypes.data, hits.ctypes.data)
    out["rays"] = rays; out["ray_hits"] = hits

    # ---- tyre model ----
    m = 400
    x = np.zeros((m, 9), np.float32)
    x[:, 0] = rng.uniform(200, 6000, m); x[:, 1] = rng.uniform(-0.6, 0.6, m); x[:, 2] = rng.uniform(-1.0, 1.0, m)
    x[:, 3] = rng.uniform(-0.08, 0.08, m); x[:, 4] = rng.uniform(0, 60, m); x[:, 5] = rng.uniform(0.7, 1.3, m)
    x[:, 6] = rng.uniform(0.05, 0.2, m); x[:, 7] = rng.uniform(-0.3, 0.3, m); x[:, 8] = rng.uniform(0, 20, m) * (rng.random(m) < 0.3)
    y = np.zeros((4, m, 7), np.float32)
    for w in range(4):
        r.L.pdref_sctm_solve(r.h, w, m, x.ctypes.data, y[w].ctypes.data)
    out["sctm_in"] = x; out["sctm_out"] = y

    # ---- collisions: whole-car translations of driving states; the flag after ONE step on an odd physics frame ----
    lay = pdref.Layout()
    bodies = ["chassis", "tank", "hub0", "strut0", "hub1", "strut1", "axle"]
    cb, cf, ct = [], [], []
    s = pdref.RefSim(); s.set_collision_response(False)
    for case in range(600):
        if case % 20 == 0:
            s.teleport_spline(float(rng.uniform(0, 1)))
            for t in range(int(rng.integers(5, 250))):
                s.set_controls(steer=float(rng.uniform(-0.3, 0.3)), gas=0.8); s.step(DT)
            base = s.state().copy(); tbase = s.time()
        rec = base.copy()
        dx, dz = rng.uniform(-9, 9, 2)
        dy = 0.0 if rng.random() < 0.4 else float(rng.uniform(-0.3, 0.02))
        if case % 5 == 0:
            dx = dz = 0.0
        for b in bodies:
            lay.set(rec, b + ".px", lay.get(rec, b + ".px") + dx); lay.set(rec, b + ".py", lay.get(rec, b + ".py") + dy); lay.set(rec, b + ".pz", lay.get(rec, b + ".pz") + dz)
        lay.set(rec, "car.physFrame", 1 + 2 * int(rng.integers(0, 50)))
        s.set_state(rec); s.set_time(tbase)
        cb.append(s.state().copy()); ct.append(tbase)
        s.step(DT)
        cf.append(lay.get(s.state(), "car.collisionFlag"))
    s.close()
    out["coll_before"] = np.stack(cb); out["coll_time"] = np.array(ct, np.float64); out["coll_flag"] = np.array(cf, np.int32)
    print("collision cases: %d of %d flagged" % (int(np.sum(cf)), len(cf)))

    path = os.path.join(HERE, "demo_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
