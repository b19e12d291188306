"""Shared helpers of the parity tests: env-like configuration of a batch and the record comparison rule.

Tolerance rule (stated once, used everywhere):
  * int32 fields (hit / lock / gear / FSM state, point ids, flags) must match EXACTLY;
  * float / double fields: |mine - ref| <= tol * max(|ref|, floor), tol = 1e-4 (north star: single tick from identical
    states), with a PER-FIELD floor in the field's unit (FLOORS below): the magnitude under which a relative error stops
    meaning anything for that quantity -- 1e-3 rad for angles, 1e-2 m for lengths, 1e-2 for slips, 5e-2 for rotation-matrix /
    quaternion entries, 0.5 m/s and 1 rad/s for body velocities (the rear axle's spin, held by five short links, moves by 5e-5 rad/s under a one-ulp nudge of the positions whatever its size), 1 N / 1 Nm / 1 K for forces, torques and temperatures.
    Components of a body's velocity / angular velocity / quaternion are judged against the norm of their vector;
    world positions (ulp 8e-6 m at 100 m) get two ulps on top.
  * a record that leaves the rule is not failed outright: it goes to the conditioning arbiter (arbitrate() below).
"""
import re
import numpy as np

import pdref

BOOKKEEPING = {"car.episodeSteps", "car.nanFlag", "car.thermalPrimed"}
_VEC_GROUPS = [("px", "py", "pz"), ("vx", "vy", "vz"), ("wx", "wy", "wz"), ("qw", "qx", "qy", "qz")]
_BODIES = ["chassis", "tank", "hub0", "strut0", "hub1", "strut1", "axle", "hub3"]


def make_env_like(batch):
    """Configure a Batch the way pyprojectd/projectd_env.py:118-136 configures its simulator."""
    batch.set_assists(True, True, True)
    batch.set_collision_response(True)
    for k, v in pdref.ENV_TUNES.items():
        batch.set_tune(k, v)
    for k, v in pdref.ENV_SCORING.items():
        batch.set_scoring_var(k, v)
    return batch


_LAYOUT_CACHE = {}


def _tables(lay):
    key = id(lay)
    if key in _LAYOUT_CACHE:
        return _LAYOUT_CACHE[key]
    names = [n for n in lay.fields if n not in BOOKKEEPING]
    f_idx = np.array([lay.fields[n][0] for n in names if lay.fields[n][1] == "F"], dtype=np.int64)
    f_names = [n for n in names if lay.fields[n][1] == "F"]
    i_idx = np.array([lay.fields[n][0] for n in names if lay.fields[n][1] == "I"], dtype=np.int64)
    i_names = [n for n in names if lay.fields[n][1] == "I"]
    d_idx = np.array([lay.fields[n][0] for n in names if lay.fields[n][1] == "D"], dtype=np.int64)
    d_names = [n for n in names if lay.fields[n][1] == "D"]
    # vector scale groups: position of f_names entries belonging to one body vector
    pos = {n: k for k, n in enumerate(f_names)}
    groups = []
    for b in _BODIES:
        for g in _VEC_GROUPS:
            groups.append([pos["%s.%s" % (b, c)] for c in g])
    t = (f_idx, f_names, i_idx, i_names, d_idx, d_names, groups)
    _LAYOUT_CACHE[key] = t
    return t


FLOORS = [   # (regex on the field name, floor); first match wins
    (r"\.(px|py|pz|contactX|contactY|contactZ)$", 1e-2), (r"car\.pointCache", 1e-2),
    (r"\.(qw|qx|qy|qz|a[xyz][xyz]|normalX|normalY|normalZ)$", 5e-2),
    (r"\.(vx|vy|vz)$", 0.5), (r"\.(wx|wy|wz)$", 1.0), (r"car\.(lastVel|accG|speed)", 1e-1),
    (r"(slipAngleRAD|camberRAD|finalSteerAngleSignal|lookAhead|currentDriftAngle)", 1e-3),
    (r"(slipRatio|ndSlip|slipFactor|dirtyLevel|wearMult|inflation|thermalMultD)", 1e-2),
    (r"(depth|distToGround|Radius|suspTravel)$", 1e-2), (r"(suspDamperSpeed|totalHubVelocity|slidingVelocity)", 1e-1),
    (r"car\.(ctl|smoothSteerValue|acClutchValueSignal|gasCutoff|bodyVsTrack|velocityVsTrack|trackLocation|oldTrackLocation|locClutch)", 1e-2),
    (r"car\.(probe|stepReward|totalReward|prevEpisodeReward|instantDrift|driftPoints|lastTrackPointTimestamp|acSeqTime|driftStraightTimer)", 1e-1),
]
_FLOOR_CACHE = {}


def field_floor(name):
    if name not in _FLOOR_CACHE:
        fl = 1.0        # forces (N), torques (Nm), temperatures, pressures, rad/s of shafts and wheels, rpm-like doubles, damage
        for pat, v in FLOORS:
            if re.search(pat, name):
                fl = v; break
        _FLOOR_CACHE[name] = fl
    return _FLOOR_CACHE[name]


_VEC_FLOOR = {"px": None, "vx": 1e-1, "wx": 1e-1}


def compare_records(lay, mine, ref, tol=1e-4, floor=None):
    """Returns (list of (field, mine, ref, rel_err) violating the rule, worst relative error over float fields).
    floor=None: the per-field table; a number: that floor for every field (the round-1 rule was floor=1.0)."""
    f_idx, f_names, i_idx, i_names, d_idx, d_names, groups = _tables(lay)
    key = ("floors", id(lay), floor)
    if key not in _LAYOUT_CACHE:
        _LAYOUT_CACHE[key] = (np.array([field_floor(n) if floor is None else floor for n in f_names]), np.array([field_floor(n) if floor is None else floor for n in d_names]),
                              np.array([bool(re.search(r"\.(px|py|pz|contactX|contactY|contactZ)$", n)) for n in f_names]))
    f_floor, d_floor, is_pos = _LAYOUT_CACHE[key]
    mine = np.ascontiguousarray(mine, dtype=np.uint32); ref = np.ascontiguousarray(ref, dtype=np.uint32)
    bad = []
    mi = mine[i_idx].view(np.int32); ri = ref[i_idx].view(np.int32)
    for k in np.nonzero(mi != ri)[0]:
        bad.append((i_names[k], int(mi[k]), int(ri[k]), float("inf")))
    mf = mine[f_idx].view(np.float32).astype(np.float64); rf = ref[f_idx].view(np.float32).astype(np.float64)
    scale = np.maximum(np.abs(rf), f_floor)
    for g in groups:
        if is_pos[g[0]]:
            continue                                  # positions: per component, floor + two ulps (below)
        scale[g] = max(float(np.sqrt((rf[g] ** 2).sum())), float(f_floor[g[0]]))
    scale[is_pos] = f_floor[is_pos] + 2.0 * np.spacing(np.abs(rf[is_pos]).astype(np.float32)).astype(np.float64) / tol
    with np.errstate(invalid="ignore"):
        rel = np.abs(mf - rf) / scale
    rel = np.where(np.isnan(rel), np.where(np.isnan(mf) & np.isnan(rf), 0.0, np.inf), rel)
    for k in np.nonzero(rel > tol)[0]:
        bad.append((f_names[k], float(mf[k]), float(rf[k]), float(rel[k])))
    worst = float(rel.max()) if len(rel) else 0.0
    if len(d_idx):
        md = np.array([mine[o:o + 2].view(np.float64)[0] for o in d_idx]); rd = np.array([ref[o:o + 2].view(np.float64)[0] for o in d_idx])
        reld = np.abs(md - rd) / np.maximum(np.abs(rd), d_floor)
        for k in np.nonzero(reld > tol)[0]:
            bad.append((d_names[k], float(md[k]), float(rd[k]), float(reld[k])))
        worst = max(worst, float(reld.max()))
    return bad, worst


def params_equal(mine, ref, hostsim_lib):
    """Byte comparison of two PdCarParams blocks.  One documented difference is tolerated: a car without [AUTO_SHIFTER] leaves
    AutoShifter::changeUpRpm / changeDnRpm at (0, 4000) until its first active step (AutoShifter.cpp:38-53), the loader resolves
    them at load time; a reference block still holding (0, 4000) accepts the loader's resolved pair."""
    mine = np.array(mine, dtype=np.uint8)[: len(ref)].copy(); ref = np.array(ref, dtype=np.uint8)
    o = hostsim_lib.hs_offset_autoshift_rpm()
    if tuple(ref[o:o + 8].view(np.int32)) == (0, 4000):
        mine[o:o + 8] = ref[o:o + 8]
    return np.array_equal(mine, ref)


def scripted_controls(t, phase=0.0, gas_scale=1.0):
    """BASELINE.json config 1 script: gas = 0.1+0.9*min(1,t/333), steer = 0.3*sin(2*pi*t/999)."""
    import math
    gas = (0.1 + 0.9 * min(1.0, t / 333.0)) * gas_scale
    steer = 0.3 * math.sin(2 * math.pi * t / 999.0 + phase)
    return steer, gas


def centering_steer(lay, rec, gain=1.5):
    """Keeps a car on the road: steer against the difference of the two 25-degree probes (state of the last tick)."""
    p1 = lay.get(rec, "car.probe1"); p2 = lay.get(rec, "car.probe2")
    return max(-1.0, min(1.0, -gain * (p1 - p2) / (p1 + p2 + 1e-3)))


def drive_controls(t, i, lay=None, rec=None):
    """Scripted drive i at tick t -> keyword arguments of RefSim.set_controls.
      kind 0 (i % 64 in 0..15)  config-1 style throttle ramp + steering sine from the grid (the round-1 drives);
      kind 1 (16..31) from a rolling start (see drive_start_states): handbrake turns and brake-to-lock at speed;
      kind 2 (32..47) from the grid: reverse-gear launch through the sequential shifter, brake to a standstill and wait
                      (sleep counter), forward again;
      kind 3 (48..63) from a rolling start: first gear held on the rev limiter, then full steering lock at speed (the car
                      leaves the tarmac: grass / sand / kerb surfaces).
    rec = the oracle's state before the tick (for the road-keeping steering of the rolling drives)."""
    steer, gas = scripted_controls(t, phase=0.7 * i, gas_scale=0.4 + 0.6 * ((i % 4) / 3.0))
    c = dict(steer=steer, gas=gas, brake=0.0, hand_brake=0.0, clutch=0.0, req_gear=-1, gear_up=0, gear_dn=0)
    kind = (i // 16) % 4
    if i % 5 == 4 and 400 < t < 450:
        c["brake"] = 0.6
    if kind in (1, 3) and rec is not None:
        c["steer"] = centering_steer(lay, rec); c["gas"] = 1.0 if lay.get(rec, "car.speed") < 22 else 0.3
    if kind == 1:
        if 60 <= t % 300 < 130:
            c["hand_brake"] = 1.0; c["steer"] = 0.8 if i % 2 else -0.8
        if 180 <= t % 300 < 250:
            c["brake"] = 1.0; c["gas"] = 0.0
    elif kind == 2:
        if t < 6:
            c["gear_dn"] = 1 if t % 2 == 0 else 0; c["gas"] = 0.0
        elif t < 180:
            c["gas"] = 0.7
        elif t < 420:
            c["gas"] = 0.0; c["brake"] = 0.9
        elif t < 430:
            c["gear_up"] = 1 if t % 2 == 0 else 0; c["gas"] = 0.0
    elif kind == 3:
        c["gas"] = 1.0
        if t < 150:
            c["req_gear"] = 2
        if t > 200 + 10 * (i % 16):
            c["steer"] = 1.0 if i % 2 else -1.0
    return c


_START_CACHE = {}
# spline points of yamanashi_short whose left verge (2.5 m beyond the boundary) is SAND (DAMPING 0.1, DIRT_ADDITIVE 1, SIN_HEIGHT 0.04)
_SAND_POINTS = {"yamanashi_short": [444, 448, 452, 456, 460, 464]}


def place_beside_track(oracle_mod, lay, r, track, pid, side=0, extra=2.5):
    """Stand the car of RefSim r on the verge: the grid pose at spline point pid, translated sideways past the track boundary
    (side 0 = left) by `extra` metres and vertically by the ground-height difference (two oracle rays)."""
    fat = np.fromfile(oracle_mod.BASE_PATH + "/content/tracks/%s/spline.cache" % track, dtype=np.float32).reshape(-1, 15)
    best = fat[pid, 0:3]; edge = fat[pid, 3:6] if side == 0 else fat[pid, 6:9]
    d = (edge - best).astype(np.float64); d[1] = 0; w = float(np.linalg.norm(d)); d /= max(w, 1e-6)
    off = d * (w + extra)
    rays = np.zeros((2, 7), np.float32); rays[:, 4] = -1; rays[:, 6] = 20
    rays[0, 0:3] = best + np.array([0, 5, 0], np.float32); rays[1, 0:3] = best + off.astype(np.float32) + np.array([0, 5, 0], np.float32)
    out = np.zeros((2, 8), np.float32); r.L.pdref_raycast(r.h, 2, rays.ctypes.data, out.ctypes.data)
    assert out[0, 0] and out[1, 0], "no ground beside spline point %d" % pid
    dy = float(out[1, 2] - out[0, 2]) + 0.02
    r.teleport_spline(pid / len(fat))
    rec = r.state().copy()
    for bd in _BODIES:
        lay.set(rec, bd + ".px", lay.get(rec, bd + ".px") + float(off[0])); lay.set(rec, bd + ".pz", lay.get(rec, bd + ".pz") + float(off[2]))
        lay.set(rec, bd + ".py", lay.get(rec, bd + ".py") + dy)
    r.set_state(rec)


def drive_start_states(oracle_mod, lay, track, n, preroll=1200, car=None):
    """Start records of the n scripted drives on `track` (cached per session): kinds 0 / 2 stand on the grid at their spline
    position; kinds 1 / 3 have been driven `preroll` ticks by the oracle under the road-keeping steering (10-16 m/s, 3rd gear)."""
    key = (track, n, preroll, car)
    kw = {"car": car} if car else {}
    if key in _START_CACHE:
        return _START_CACHE[key]
    out = []
    for i in range(n):
        r = oracle_mod.RefSim(track=track, **kw)
        r.teleport_spline((i % 16) / 16 + (i // 16) * 0.013)
        sand = _SAND_POINTS.get(track, [])
        if (i // 16) % 4 == 3 and (i % 16) < len(sand):
            place_beside_track(oracle_mod, lay, r, track, sand[i % 16])          # these drives start on the verge
        elif (i // 16) % 4 in (1, 3):
            for t in range(preroll):
                rec = r.state()
                r.set_controls(steer=centering_steer(lay, rec), gas=1.0 if lay.get(rec, "car.speed") < 22 else 0.3)
                r.step()
        out.append((r.state().copy(), r.time(), r.frame()))
        r.close()
    _START_CACHE[key] = out
    return out


def coverage_marks(lay, rec, seen):
    """Which of the rarely visited regions of the state space the (oracle) record `rec` is in."""
    g = lambda n: lay.get(rec, n)
    gear = g("car.currentGear")
    if gear == 0 and abs(g("car.speed")) > 1.0:
        seen.add("reverse")
    if gear >= 3:
        seen.add("gear>=3")
    if g("car.ctlHandBrake") > 0.5 and g("car.speed") > 3.0:
        seen.add("handbrake")
    if g("car.limiterOn"):
        seen.add("limiter")
    if g("car.sleepingFrames") > 50:
        seen.add("sleeping")
    if g("car.isGearGrinding"):
        seen.add("grinding")
    for w in range(4):
        if g("tyre%d.isLocked" % w) and g("car.speed") > 3.0:
            seen.add("locked_wheel_at_speed")
        s = g("tyre%d.surfaceId" % w)
        if s > 0 and g("tyre%d.hasContact" % w):
            seen.add("surface%d" % s)
        if not g("tyre%d.hasContact" % w):
            seen.add("wheel_airborne")
        if g("tyre%d.dirtyLevel" % w) > 0:
            seen.add("dirty_tyre")
    return seen


# ---- conditioning arbiter -------------------------------------------------------------------------------------------------
# A field may leave the 1e-4 rule without anything being wrong: the reference's own formulation is ill-conditioned in places
# (tyre load = 254 573 N/m x a contact depth formed from world coordinates whose fp32 ulp is 4e-6 m at 40 m altitude: one ulp
# = 1 N; the rear axle's spin about its own axis is held by five links with short lever arms).  The arbiter measures that
# conditioning instead of guessing it: it re-runs THE ORACLE from the same state with every body position nudged by one fp32
# ulp and records how far each field of the oracle's own result moves.  A deviation is accepted when it stays within 4x that
# spread (+ the plain rule); anything larger is a real difference.  It is only consulted for records that failed the plain rule.
_ARB_SIMS = {}


def _nudged(lay, rec, mode):
    out = rec.copy()
    for b in _BODIES:
        for c in ("px", "py", "pz"):
            if mode == 2 and c != "py":
                continue
            off = lay.fields["%s.%s" % (b, c)][0]
            v = out[off:off + 1].view(np.float32)
            v[0] = np.nextafter(v[0], np.float32(np.inf if mode != 1 else -np.inf))
    return out


def arbitrate(oracle_mod, lay, track, before, time_before, ref_after, bad, tol=1e-4, factor=4.0, dt=1.0 / 333.0, car=None, base=None):
    """bad: list of (field, mine, ref, rel) from compare_records.  Returns the entries that remain bad after arbitration."""
    r = _ARB_SIMS.get((track, car, base))
    if r is None:
        r = _ARB_SIMS[(track, car, base)] = oracle_mod.RefSim(track=track, **({"car": car} if car else {}), **({"base": base} if base else {}))
    spread = {}
    for mode in (0, 1, 2):
        r.set_state(_nudged(lay, before, mode)); r.set_time(time_before); r.step(dt)
        alt = r.state()
        for name, mine, ref, rel in bad:
            if not np.isfinite(rel):
                continue
            spread[name] = max(spread.get(name, 0.0), abs(lay.get(alt, name) - lay.get(ref_after, name)))
    left = []
    for name, mine, ref, rel in bad:
        if not np.isfinite(rel) or abs(mine - ref) > factor * spread.get(name, 0.0) + tol * max(abs(ref), field_floor(name)):
            left.append((name, mine, ref, rel, spread.get(name, 0.0)))
    return left


def make_brake_temps_base(tmp_dir, base, car="ks_toyota_ae86_drift", new_name="ae86_brake_temps"):
    """A content tree (symlinks into `base`) with one extra car: `car` plus [TEMPS_FRONT] / [TEMPS_REAR] / [EBB] in its brakes.ini
    -- the optional BrakeSystem parts no bundled car ships data for (BrakeSystem.cpp:28-54: 'cars: F40').  Returns (base, car)."""
    import os, shutil
    root = os.path.join(str(tmp_dir), "base")
    os.makedirs(os.path.join(root, "content", "cars"), exist_ok=True)
    for d in ("cfg",):
        if not os.path.exists(os.path.join(root, d)):
            os.symlink(os.path.join(base, d), os.path.join(root, d))
    if not os.path.exists(os.path.join(root, "content", "tracks")):
        os.symlink(os.path.join(base, "content", "tracks"), os.path.join(root, "content", "tracks"))
    dst = os.path.join(root, "content", "cars", new_name)
    if not os.path.isdir(dst):
        shutil.copytree(os.path.join(base, "content", "cars", car), dst)
        with open(os.path.join(dst, "data", "brakes.ini"), "a") as f:
            f.write("\n[EBB]\nFRONT_SHARE_MULTIPLIER=1.25\n\n"
                    "[TEMPS_FRONT]\nPERF_CURVE=(|0=0.85|150=0.95|300=1.0|600=1.0|800=0.7|)\nTORQUE_K=0.12\nCOOL_TRANSFER=0.006\nCOOL_SPEED_FACTOR=0.0015\n\n"
                    "[TEMPS_REAR]\nPERF_CURVE=(|0=0.8|200=1.0|500=1.0|700=0.75|)\nTORQUE_K=0.09\nCOOL_TRANSFER=0.004\nCOOL_SPEED_FACTOR=0.001\n")
    return root, new_name


def make_variant_car_base(tmp_dir, base, kind, car="ks_mazda_rx7_tuned"):
    """A content tree (symlinks into `base`) with one car derived from the double-wishbone `car` for the suspension parts no bundled
    car ships data for.  kind "ml": the REAR axle becomes TYPE=ML (SuspensionML: JOINTn_CAR / JOINTn_TYRE taken from the wishbone
    points, so the geometry stays a sane one); kind "heave": [HEAVE_FRONT] / [HEAVE_REAR] third springs are added; kind "throttle": engine.ini gets a
    [THROTTLE_RESPONSE] section (second throttle map blended in with rpm).  Returns (base, car)."""
    import os, re, shutil
    root = os.path.join(str(tmp_dir), "base_" + kind)
    os.makedirs(os.path.join(root, "content", "cars"), exist_ok=True)
    if not os.path.exists(os.path.join(root, "cfg")):
        os.symlink(os.path.join(base, "cfg"), os.path.join(root, "cfg"))
    if not os.path.exists(os.path.join(root, "content", "tracks")):
        os.symlink(os.path.join(base, "content", "tracks"), os.path.join(root, "content", "tracks"))
    new_name = "rx7_" + kind
    dst = os.path.join(root, "content", "cars", new_name)
    if os.path.isdir(dst):
        return root, new_name
    shutil.copytree(os.path.join(base, "content", "cars", car), dst)
    path = os.path.join(dst, "data", "suspensions.ini")
    text = open(path, newline="").read()
    nl = "\r\n" if "\r\n" in text else "\n"
    if kind == "heave":
        text += nl + nl.join(["[HEAVE_FRONT]", "BUMPSTOP_UP=0.05", "BUMPSTOP_DN=0.04", "ROD_LENGTH=0.0", "SPRING_RATE=30000", "PROGRESSIVE_SPRING_RATE=2000", "DAMP_BUMP=1500", "DAMP_REBOUND=2500",
                              "DAMP_FAST_BUMP=900", "DAMP_FAST_REBOUND=1400", "DAMP_FAST_BUMPTHRESHOLD=0.1", "DAMP_FAST_REBOUNDTHRESHOLD=0.1", "BUMP_STOP_RATE=0", "PACKER_RANGE=0.06", "",
                              "[HEAVE_REAR]", "BUMPSTOP_UP=0.06", "BUMPSTOP_DN=0.03", "ROD_LENGTH=0.0", "SPRING_RATE=25000", "PROGRESSIVE_SPRING_RATE=0", "DAMP_BUMP=1200", "DAMP_REBOUND=2000",
                              "DAMP_FAST_BUMP=0", "DAMP_FAST_REBOUND=0", "DAMP_FAST_BUMPTHRESHOLD=0", "DAMP_FAST_REBOUNDTHRESHOLD=0", "BUMP_STOP_RATE=120000", "PACKER_RANGE=0.08", ""]) + nl
    elif kind == "ml":
        m = re.search(r"\[REAR\](.*?)(?=\r?\n\[|\Z)", text, re.S)
        sec = m.group(1)
        get = lambda k: re.search(r"^%s=(.*?)\s*$" % k, sec, re.M).group(1)
        joints = [("WBCAR_TOP_REAR", "WBTYRE_TOP"), ("WBCAR_TOP_FRONT", "WBTYRE_TOP"), ("WBCAR_BOTTOM_REAR", "WBTYRE_BOTTOM"), ("WBCAR_BOTTOM_FRONT", "WBTYRE_BOTTOM"), ("WBCAR_STEER", "WBTYRE_STEER")]
        extra = nl.join("JOINT%d_CAR=%s%sJOINT%d_TYRE=%s" % (i, get(a), nl, i, get(b)) for i, (a, b) in enumerate(joints))
        sec2 = re.sub(r"^TYPE=DWB", "TYPE=ML", sec, flags=re.M).rstrip() + nl + extra + nl
        text = text[:m.start(1)] + sec2 + text[m.end(1):]
    elif kind == "throttle":
        epath = os.path.join(dst, "data", "engine.ini")
        etext = open(epath, newline="").read()
        enl = "\r\n" if "\r\n" in etext else "\n"
        open(epath, "w", newline="").write(etext + enl + enl.join(["[THROTTLE_RESPONSE]", "RPM_REFERENCE=6000", "LUT=(|0=0|30=55|60=85|100=100|)", ""]) + enl)
    else:
        raise ValueError(kind)
    open(path, "w", newline="").write(text)
    return root, new_name
