"""Drop-in mirror of the reference's ``PyProjectD`` module (src/PyProjectD/PyProjectD.cpp:515-641) for the
Car::step hot path, on top of the CUDA batch (``libpd_b200.so``).

Same function names, argument meaning and error behaviour as the reference:

* nothing raises across this boundary: failures are logged and give the sentinel ``-1`` (ids) or a silent
  no-op on bad ids, as PyProjectD.cpp:100-136 does;
* a *simulator* (``simId``) is one CUDA batch; ``createSimulator(basePath, numEnvs=1, device=0)`` has two extra
  optional arguments, with the defaults it is the reference's call.  The reference's usage is one car per
  simulator (pyprojectd/projectd_env.py:118-121); here ``carId`` is the env index inside the batch
  (``addCar`` returns 0; ids ``0..numEnvs-1`` are valid and share the model).
* batched variants (suffix ``Batch``) take / return arrays for all envs and ``getObsTensor`` hands out the
  observation matrix [numEnvs, 24] through DLPack without a copy.

Viewer / playground functions of the reference (initPlayground ...) are outside the hot path and absent.
"""
from __future__ import annotations

import sys
import threading

import numpy as np

from .binding import Batch, PdError

_sims: dict[int, "_Sim"] = {}
_mux = threading.Lock()
_next_id = 0
_seed = 0
_log_file = None


class vec3f:
    """Core/Math.h vec3f (x, y, z read-write)."""
    __slots__ = ("x", "y", "z")

    def __init__(self, x=0.0, y=0.0, z=0.0):
        self.x, self.y, self.z = float(x), float(y), float(z)

    def __iter__(self):
        return iter((self.x, self.y, self.z))

    def __repr__(self):
        return "vec3f(%g, %g, %g)" % (self.x, self.y, self.z)


class mat44f:
    """Core/Math.h mat44f, row-vector convention: M11..M13 = body x axis in world, M41..M43 = translation."""
    _names = ["M%d%d" % (r, c) for r in range(1, 5) for c in range(1, 5)]
    __slots__ = _names

    def __init__(self, values=None):
        v = list(values) if values is not None else [0.0] * 16
        for n, x in zip(self._names, v):
            setattr(self, n, float(x))


class CarControls:
    """Car/CarControls.h:9-20."""
    __slots__ = ("steer", "clutch", "brake", "handBrake", "gas", "isShifterSupported", "requestedGearIndex", "gearUp", "gearDn")

    def __init__(self):
        self.steer = 0.0; self.clutch = 0.0; self.brake = 0.0; self.handBrake = 0.0; self.gas = 0.0
        self.isShifterSupported = 1; self.requestedGearIndex = -1; self.gearUp = 0; self.gearDn = 0


_CTRL_DTYPE = np.dtype([("steer", "<f4"), ("clutch", "<f4"), ("brake", "<f4"), ("handBrake", "<f4"), ("gas", "<f4"),
                        ("isShifterSupported", "i1"), ("requestedGearIndex", "i1"), ("gearUp", "i1"), ("gearDn", "i1")])
#: Car/CarState.h:11-56, pack(4), 664 bytes
CAR_STATE_DTYPE = np.dtype([
    ("carId", "<i4"), ("simId", "<i4"), ("timestamp", "<f4"), ("controls", _CTRL_DTYPE),
    ("collisionFlag", "<i4"), ("outOfTrackFlag", "<i4"), ("trackPointId", "<i4"), ("lastTrackPointTimestamp", "<f4"),
    ("trackLocation", "<f4"), ("bodyVsTrack", "<f4"), ("velocityVsTrack", "<f4"),
    ("engineRPM", "<f4"), ("speedMS", "<f4"), ("gear", "<i4"), ("gearGrinding", "<i4"),
    ("bodyMatrix", "<f4", (16,)), ("bodyPos", "<f4", (3,)), ("bodyEuler", "<f4", (3,)), ("accG", "<f4", (3,)),
    ("velocity", "<f4", (3,)), ("localVelocity", "<f4", (3,)), ("angularVelocity", "<f4", (3,)), ("localAngularVelocity", "<f4", (3,)),
    ("hubMatrix", "<f4", (4, 16)), ("tyreContacts", "<f4", (4, 3)), ("tyreLoad", "<f4", (4,)), ("tyreAngularSpeed", "<f4", (4,)),
    ("tyreSlipRatio", "<f4", (4,)), ("tyreNdSlip", "<f4", (4,)),
    ("probes", "<f4", (10,)), ("lookAhead", "<f4", (5,)), ("stepReward", "<f4"), ("totalReward", "<f4"),
])
assert CAR_STATE_DTYPE.itemsize == 664

_VEC_FIELDS = ("bodyPos", "bodyEuler", "accG", "velocity", "localVelocity", "angularVelocity", "localAngularVelocity")


class CarState:
    """Car/CarState.h; the attribute shapes follow the pybind11 bindings (vec3f / mat44f objects, lists)."""

    def __init__(self):
        self._fill(np.zeros((), CAR_STATE_DTYPE))

    def _fill(self, rec):
        for name in CAR_STATE_DTYPE.names:
            v = rec[name]
            if name == "controls":
                c = CarControls()
                for n in _CTRL_DTYPE.names:
                    setattr(c, n, v[n].item())
                v = c
            elif name in _VEC_FIELDS:
                v = vec3f(*v.tolist())
            elif name == "bodyMatrix":
                v = mat44f(v.tolist())
            elif name == "hubMatrix":
                v = [mat44f(m.tolist()) for m in v]
            elif name == "tyreContacts":
                v = [vec3f(*m.tolist()) for m in v]
            elif v.ndim:
                v = v.tolist()
            else:
                v = v.item()
            object.__setattr__(self, name, v)


class _Sim:
    def __init__(self, base, n_envs, device):
        self.base, self.n_envs, self.device = base, n_envs, device
        self.track = None
        self.batch: Batch | None = None
        self.auto_teleport = (False, False, 0)


def _log(msg):
    line = "[PyProjectD/b200] " + msg
    if _log_file:
        try:
            with open(_log_file, "a") as f:
                f.write(line + "\n")
            return
        except OSError:
            pass
    print(line, file=sys.stderr)


def _batch(simId, carId=0) -> Batch | None:
    with _mux:
        s = _sims.get(simId)
    if s is None or s.batch is None or not (0 <= carId < s.n_envs):
        return None
    return s.batch


def _guard(fn):
    def wrapped(*a, **k):
        try:
            return fn(*a, **k)
        except Exception as ex:     # nothing crosses the boundary (PyProjectD.cpp:132-136)
            _log("%s: %s" % (fn.__name__, ex))
            return -1 if fn.__name__ in ("createSimulator", "addCar") else None
    wrapped.__name__ = fn.__name__; wrapped.__doc__ = fn.__doc__
    return wrapped


# ---------------------------------------------------------------- module functions (reference names) ----------
def setSeed(seed):
    """PyProjectD.cpp:50-53.  Seeds the counter-based env RNG (random teleports are keyed by seed + env id)."""
    global _seed
    _seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    with _mux:
        sims = list(_sims.values())
    for s in sims:
        if s.batch is not None:
            s.batch.set_seed(_seed, 0)


def setLogFile(path, overwrite=False):
    global _log_file
    _log_file = path
    if overwrite:
        open(path, "w").close()


def clearLogFile():
    if _log_file:
        open(_log_file, "w").close()


def writeLog(msg):
    _log(str(msg))


@_guard
def createSimulator(basePath, numEnvs=1, device=0):
    """PyProjectD.cpp:111-137: returns the simulator id (>= 0) or -1."""
    global _next_id
    if numEnvs <= 0:
        raise ValueError("numEnvs must be positive")
    with _mux:
        sid = _next_id; _next_id += 1
        _sims[sid] = _Sim(str(basePath), int(numEnvs), int(device))
    return sid


@_guard
def destroySimulator(simId):
    with _mux:
        s = _sims.pop(simId, None)
    if s is not None and s.batch is not None:
        s.batch.close()


@_guard
def loadTrack(simId, trackName):
    with _mux:
        s = _sims.get(simId)
    if s is not None:
        s.track = str(trackName)


@_guard
def unloadTrack(simId):
    with _mux:
        s = _sims.get(simId)
    if s is not None:
        if s.batch is not None:
            s.batch.close(); s.batch = None
        s.track = None


@_guard
def addCar(simId, modelName):
    """PyProjectD.cpp:219-243: car id, -1 on failure.  Creates the CUDA batch (track must be loaded first, as the
    reference requires: Car::init places the car on the track)."""
    with _mux:
        s = _sims.get(simId)
    if s is None or s.track is None:
        _log("addCar: no such simulator or no track loaded")
        return -1
    if s.batch is not None:
        _log("addCar: this build holds one car model per simulator")
        return -1
    s.batch = Batch(s.base, track=s.track, car=str(modelName), n_envs=s.n_envs, device=s.device)
    s.batch.set_seed(_seed, 0)
    return 0


@_guard
def removeCar(simId, carId):
    with _mux:
        s = _sims.get(simId)
    if s is not None and s.batch is not None:
        s.batch.close(); s.batch = None


def _mask(n, carId):
    m = np.zeros(n, np.uint8); m[carId] = 1
    return m


@_guard
def teleportCarToSpline(simId, carId, distanceNorm):
    b = _batch(simId, carId)
    if b is not None:
        u = np.zeros(b.n, np.float32); u[carId] = distanceNorm
        b.teleport_spline(u, _mask(b.n, carId))


@_guard
def teleportCarByMode(simId, carId, mode):
    """0 start / 1 nearest / 2 random (Car.cpp:1342-1358)."""
    b = _batch(simId, carId)
    if b is not None:
        b.teleport_mode(int(mode), _mask(b.n, carId))


@_guard
def setCarAutoTeleport(simId, carId, onCollision, onBadLocation, mode=0):
    """Reference: teleports inside ScoringSystem when the flag fires.  Here the batched env step owns resets
    (pd_env_step); the setting is recorded and honoured by stepSimulator below."""
    with _mux:
        s = _sims.get(simId)
    if s is not None:
        s.auto_teleport = (bool(onCollision), bool(onBadLocation), int(mode))


@_guard
def setCarAssists(simId, carId, autoClutch, autoShift, autoBlip):
    b = _batch(simId, carId)
    if b is not None:
        b.set_assists(bool(autoClutch), bool(autoShift), bool(autoBlip))


@_guard
def setCarTune(simId, carId, name, value):
    b = _batch(simId, carId)
    if b is not None:
        b.set_tune(str(name), float(value))


@_guard
def setCarRawTune(simId, carId, name, value):
    """PyProjectD.cpp:337-344 -> SetupManager::setRawTune: the raw value, no spinner clamp / multiplier."""
    b = _batch(simId, carId)
    if b is not None:
        b.set_raw_tune(str(name), float(value))


@_guard
def setScoringVar(simId, carId, name, value):
    b = _batch(simId, carId)
    if b is not None:
        b.set_scoring_var(str(name), float(value))


@_guard
def getScoringVar(simId, carId, name):
    b = _batch(simId, carId)
    return b.get_scoring_var(str(name)) if b is not None else 0.0


@_guard
def setCarControls(simId, carId, smooth, controls):
    """PyProjectD.cpp:297-305 for one env of the batch (the other envs keep their controls)."""
    b = _batch(simId, carId)
    if b is None:
        return
    ctl, gears = b.controls_host()
    ctl[carId] = (controls.steer, controls.clutch, controls.brake, controls.handBrake, controls.gas)
    gears[carId] = (controls.requestedGearIndex, controls.gearUp, controls.gearDn)     # the struct is copied as is (PyProjectD.cpp:297-305)
    b.set_controls(ctl, gears, smooth=bool(smooth))


@_guard
def setCarControlsBatch(simId, controls, gears=None, smooth=True):
    """controls [numEnvs,5] = steer, clutch, brake, handBrake, gas; gears [numEnvs,3] int8 = requestedGearIndex
    (-1 sequential), gearUp, gearDn."""
    b = _batch(simId)
    if b is not None:
        b.set_controls(controls, gears, smooth=bool(smooth))


@_guard
def stepSimulator(simId, dt):
    """Simulator::step (Sim/Simulator.cpp:168-200) for every env of the batch."""
    with _mux:
        s = _sims.get(simId)
    if s is None or s.batch is None:
        return
    s.batch.step(float(dt), 1)
    on_col, on_bad, mode = s.auto_teleport
    if on_col or on_bad:
        _, _, flags = s.batch.rewards()
        m = np.zeros(s.n_envs, np.uint8)
        if on_col:
            m |= (flags & 1).astype(np.uint8)
        if on_bad:
            m |= ((flags >> 1) & 1).astype(np.uint8)
        if m.any():
            s.batch.teleport_mode(mode, m)


@_guard
def getCarState(simId, carId, state):
    """Fills the caller's CarState (PyProjectD.cpp:307-316), 664-byte layout of Car/CarState.h."""
    b = _batch(simId, carId)
    if b is None:
        return
    rec = np.frombuffer(b.car_state_bytes(carId), dtype=CAR_STATE_DTYPE)[0]
    state._fill(rec)
    state.simId = simId; state.carId = carId


@_guard
def getObsTensor(simId):
    """Observation matrix [numEnvs, 24] f32 on the GPU (obs selection of projectd_env.py:237-275), via DLPack."""
    b = _batch(simId)
    if b is None:
        return None
    b.observe()
    return b.obs_tensor()


def getBatch(simId):
    """The underlying :class:`projectd_core_b200.Batch` (env stepping, snapshots, raycasts)."""
    return _batch(simId)


def shutAll():
    with _mux:
        ids = list(_sims)
    for i in ids:
        destroySimulator(i)
