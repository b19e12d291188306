"""Shared helpers of the parity tests: env-like configuration of a batch and the record comparison rule.

Tolerance rule (stated once, used everywhere):
  * int32 fields (hit / lock / gear / FSM state, point ids, flags) must match EXACTLY;
  * float / double fields: |mine - ref| <= tol * max(|ref|, floor) with floor = 1.0 in the field's unit
    (1 m, 1 m/s, 1 rad/s, 1 N ...).  For the components of a rigid body's position / velocity / angular
    velocity the scale is the norm of that body's vector, so a tiny component of a large vector is judged against
    the vector.  The north-star tolerance is tol = 1e-4 for one tick from identical states.
"""
import numpy as np

import pdref

BOOKKEEPING = {"car.episodeSteps", "car.nanFlag", "car.thermalPrimed"}
_VEC_GROUPS = [("px", "py", "pz"), ("vx", "vy", "vz"), ("wx", "wy", "wz")]
_BODIES = ["chassis", "tank", "hub0", "strut0", "hub1", "strut1", "axle"]


def make_env_like(batch):
    """Configure a Batch the way pyprojectd/projectd_env.py:118-136 configures its simulator."""
    batch.set_assists(True, True, True)
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


def compare_records(lay, mine, ref, tol=1e-4, floor=1.0):
    """Returns (list of (field, mine, ref, rel_err) violating the rule, worst relative error over float fields)."""
    f_idx, f_names, i_idx, i_names, d_idx, d_names, groups = _tables(lay)
    mine = np.ascontiguousarray(mine, dtype=np.uint32); ref = np.ascontiguousarray(ref, dtype=np.uint32)
    bad = []
    mi = mine[i_idx].view(np.int32); ri = ref[i_idx].view(np.int32)
    for k in np.nonzero(mi != ri)[0]:
        bad.append((i_names[k], int(mi[k]), int(ri[k]), float("inf")))
    mf = mine[f_idx].view(np.float32).astype(np.float64); rf = ref[f_idx].view(np.float32).astype(np.float64)
    scale = np.maximum(np.abs(rf), floor)
    for g in groups:
        nrm = max(float(np.sqrt((rf[g] ** 2).sum())), floor)
        scale[g] = nrm
    with np.errstate(invalid="ignore"):
        rel = np.abs(mf - rf) / scale
    rel = np.where(np.isnan(rel), np.where(np.isnan(mf) & np.isnan(rf), 0.0, np.inf), rel)
    for k in np.nonzero(rel > tol)[0]:
        bad.append((f_names[k], float(mf[k]), float(rf[k]), float(rel[k])))
    worst = float(rel.max()) if len(rel) else 0.0
    if len(d_idx):
        md = np.array([mine[o:o + 2].view(np.float64)[0] for o in d_idx]); rd = np.array([ref[o:o + 2].view(np.float64)[0] for o in d_idx])
        reld = np.abs(md - rd) / np.maximum(np.abs(rd), floor)
        for k in np.nonzero(reld > tol)[0]:
            bad.append((d_names[k], float(md[k]), float(rd[k]), float(reld[k])))
        worst = max(worst, float(reld.max()))
    return bad, worst


def scripted_controls(t, phase=0.0, gas_scale=1.0):
    """BASELINE.json config 1 script: gas = 0.1+0.9*min(1,t/333), steer = 0.3*sin(2*pi*t/999)."""
    import math
    gas = (0.1 + 0.9 * min(1.0, t / 333.0)) * gas_scale
    steer = 0.3 * math.sin(2 * math.pi * t / 999.0 + phase)
    return steer, gas
