"""ctypes wrapper around oracle/_ref/libpdref.so (TEST INFRASTRUCTURE: the compiled reference + restated ODE).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this."""
import ctypes
import os
import re
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))   # oracle/ sits at the repo root
LIB_PATH = os.path.join(ROOT, "oracle", "_ref", "libpdref.so")
BASE_PATH = os.path.join(ROOT, "assets", "_base")   # content copy made by `make -C oracle content` (git-ignored, travels to the GPU box)

ENV_TUNES = {  # pyprojectd/projectd_env.py:83-94
    "FRONT_BIAS": 55.0, "DIFF_POWER": 30.0, "DIFF_COAST": 30.0, "FINAL_RATIO": 5.0,
    "PRESSURE_LF": 28.0, "PRESSURE_RF": 28.0, "PRESSURE_LR": 28.0, "PRESSURE_RR": 28.0,
}
ENV_SCORING = {  # pyprojectd/projectd_env.py:57-81 (only the entries that differ from ScoringSystem.cpp:50-73 matter)
    "SmoothSteerSpeed": 10.0, "MinBonusSpeed": 5.0, "MaxBonusSpeed": 200.0, "StallRpm": 300.0,
    "DirectionThreshold": 0.75, "OutOfTrackThreshold": 0.51, "ApproachDistance": 3.5, "CriticalDistance": 2.0,
    "TravelBonus": 0.1, "TravelSplineBonus": 0.01, "DriftBonus": 0.0, "SpeedBonus": 0.0, "ThrottleBonus": 0.0,
    "EngineRpmBonus": 0.0, "DirectionBonus": 0.0, "DirectionPenalty": 0.0, "ObstApproachPenalty": 0.0,
    "CollisionPenalty": 0.0, "OffTrackPenalty": 0.0, "GearGrindPenalty": 0.0, "StallPenalty": 0.0,
}


def available():
    return os.path.exists(LIB_PATH) and os.path.isdir(BASE_PATH)


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(LIB_PATH)
        vp, cp, f, i, d = ctypes.c_void_p, ctypes.c_char_p, ctypes.c_float, ctypes.c_int, ctypes.c_double
        L.pdref_create.restype = vp; L.pdref_create.argtypes = [cp, cp, cp]
        L.pdref_destroy.argtypes = [vp]
        L.pdref_teleport_spline.argtypes = [vp, f]
        L.pdref_teleport_mode.argtypes = [vp, i]
        L.pdref_set_seed.argtypes = [ctypes.c_uint]
        L.pdref_set_auto_teleport.argtypes = [vp, i, i, i]
        L.pdref_set_assists.argtypes = [vp, i, i, i]
        L.pdref_set_tune.argtypes = [vp, cp, f]
        L.pdref_set_scoring_var.argtypes = [vp, cp, f]
        L.pdref_set_controls.argtypes = [vp, f, f, f, f, f, i, i, i, i]
        L.pdref_step.argtypes = [vp, d]
        L.pdref_get_time.restype = d; L.pdref_get_time.argtypes = [vp]
        L.pdref_set_time.argtypes = [vp, d]
        L.pdref_get_state.argtypes = [vp, vp]
        L.pdref_set_state.argtypes = [vp, vp]
        L.pdref_get_car_state.argtypes = [vp, vp]
        L.pdref_get_params.argtypes = [vp, vp]
        L.pdref_get_track_info.argtypes = [vp, vp]
        L.pdref_get_spline_nodes.argtypes = [vp, vp, vp]
        L.pdref_sctm_solve.argtypes = [vp, i, i, vp, vp]
        L.pdref_raycast.argtypes = [vp, i, vp, vp]
        L.pdref_num_rows.argtypes = [vp]
        L.pdref_contacts.argtypes = [vp, vp, i]
        L.pdref_set_collision_response.argtypes = [vp, i]
        L.pdref_frame.restype = ctypes.c_uint; L.pdref_frame.argtypes = [vp]
        L.pdref_set_frame_counter.argtypes = [vp, ctypes.c_uint]
        _lib = L
    return _lib


class RefSim:
    """One reference simulator with one car, configured like pyprojectd/projectd_env.py:118-136."""

    def __init__(self, track="driftplayground", car="ks_toyota_ae86_drift", env_setup=True, base=None):
        L = lib()
        self.L = L
        self.h = L.pdref_create((base or BASE_PATH).encode(), track.encode(), car.encode())
        if not self.h:
            raise RuntimeError("pdref_create failed")
        self.words = L.pdref_state_words()
        if env_setup:
            L.pdref_teleport_mode(self.h, 0)
            L.pdref_set_auto_teleport(self.h, 0, 0, 0)
            L.pdref_set_assists(self.h, 1, 1, 1)
            for k, v in ENV_TUNES.items():
                L.pdref_set_tune(self.h, k.encode(), v)
            for k, v in ENV_SCORING.items():
                L.pdref_set_scoring_var(self.h, k.encode(), v)

    def close(self):
        if self.h:
            self.L.pdref_destroy(self.h)
            self.h = None

    def set_controls(self, steer=0.0, gas=0.0, brake=0.0, clutch=0.0, hand_brake=0.0, req_gear=-1, gear_up=0, gear_dn=0, smooth=1):
        self.L.pdref_set_controls(self.h, steer, clutch, brake, hand_brake, gas, req_gear, gear_up, gear_dn, smooth)

    def step(self, dt=1.0 / 333.0):
        self.L.pdref_step(self.h, dt)

    def state(self):
        rec = np.zeros(self.words, dtype=np.uint32)
        self.L.pdref_get_state(self.h, rec.ctypes.data)
        return rec

    def set_state(self, rec):
        rec = np.ascontiguousarray(rec, dtype=np.uint32)
        self.L.pdref_set_state(self.h, rec.ctypes.data)

    def time(self):
        return self.L.pdref_get_time(self.h)

    def set_time(self, t):
        self.L.pdref_set_time(self.h, t)

    def contacts(self):
        """Contact joints alive after the last step: array [k, 8] = position, normal, depth, kind (0 floor / 1 wall)."""
        buf = np.zeros((16, 8), np.float32)
        k = self.L.pdref_contacts(self.h, buf.ctypes.data, 16)
        return buf[:min(k, 16)]

    def set_collision_response(self, on=True):
        self.L.pdref_set_collision_response(self.h, int(bool(on)))

    def frame(self):
        """PhysicsEngineODE::currentFrame: collisions against the static meshes are tested on odd frames only."""
        return int(self.L.pdref_frame(self.h))

    def set_frame(self, f):
        self.L.pdref_set_frame_counter(self.h, int(f))

    def teleport_spline(self, u):
        self.L.pdref_teleport_spline(self.h, u)

    def params_bytes(self):
        n = self.L.pdref_params_bytes()
        buf = np.zeros(n, dtype=np.uint8)
        self.L.pdref_get_params(self.h, buf.ctypes.data)
        return buf


# ---- record layout parsed from include/pd_state.h (single source of truth) ----
def _parse_fields(text, macro):
    m = re.search(r"#define %s\(X\)(.*?)\n\n" % macro, text, re.S)
    body = re.sub(r"/\*.*?\*/", "", m.group(1), flags=re.S)
    return re.findall(r"X\((\w),\s*(\w+)\)", body)


class Layout:
    def __init__(self):
        text = open(os.path.join(ROOT, "include", "pd_state.h")).read()
        W = {"F": 1, "I": 1, "D": 2}
        self.body = _parse_fields(text, "PD_BODY_FIELDS")
        self.tyre = _parse_fields(text, "PD_TYRE_FIELDS")
        self.car = _parse_fields(text, "PD_CAR_FIELDS")
        self.fields = {}  # name -> (offset, kind)
        off = 0
        bnames = ["chassis", "tank", "hub0", "strut0", "hub1", "strut1", "axle", "hub3"]      # slot 6 = rigid axle or the LR hub of a DWB rear axle, slot 7 = its RR hub
        for b in bnames:
            for k, n in self.body:
                self.fields["%s.%s" % (b, n)] = (off, k); off += W[k]
        for w in range(4):
            for k, n in self.tyre:
                self.fields["tyre%d.%s" % (w, n)] = (off, k); off += W[k]
            for p in range(36):
                self.fields["tyre%d.T%d" % (w, p)] = (off, "F"); off += 1
        for k, n in self.car:
            self.fields["car.%s" % n] = (off, k); off += W[k]
        for p in range(10):
            self.fields["car.probe%d" % p] = (off, "F"); off += 1
        for p in range(5):
            self.fields["car.lookAhead%d" % p] = (off, "F"); off += 1
        self.words = off

    def get(self, rec, name):
        off, k = self.fields[name]
        if k == "F":
            return float(rec[off:off + 1].view(np.float32)[0])
        if k == "I":
            return int(rec[off:off + 1].view(np.int32)[0])
        return float(rec[off:off + 2].view(np.float64)[0])

    def set(self, rec, name, value):
        off, k = self.fields[name]
        if k == "F":
            rec[off:off + 1] = np.array([value], dtype=np.float32).view(np.uint32)
        elif k == "I":
            rec[off:off + 1] = np.array([value], dtype=np.int32).view(np.uint32)
        else:
            rec[off:off + 2] = np.array([value], dtype=np.float64).view(np.uint32)

    def as_dict(self, rec):
        return {n: self.get(rec, n) for n in self.fields}
