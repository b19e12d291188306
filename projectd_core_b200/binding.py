"""ctypes binding of the C ABI in include/pd_batch.h (libpd_b200.so).

The library is the product: if it is missing or no CUDA device is usable, :class:`Batch` raises -- nothing here
falls back to a CPU implementation.
"""
import ctypes
import os
import re

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
_LIB = None


class PdError(RuntimeError):
    pass


def lib_path():
    """The in-tree CUDA library; PD_B200_LIB selects another build of it (kernel A/B experiments)."""
    return os.environ.get("PD_B200_LIB") or os.path.join(_HERE, "libpd_b200.so")


def _header_constant(name):
    text = open(os.path.join(_ROOT, "include", "pd_state.h")).read()
    m = re.search(r"#define\s+%s\s+(\d+)" % name, text)
    return int(m.group(1))


OBS_DIM = 24


def load_library():
    """Load libpd_b200.so (raises PdError if it has not been built: run __graft_entry__.build())."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise PdError("CUDA library %s not built (python -c 'import __graft_entry__ as g; g.build()'); there is no CPU fallback" % path)
    L = ctypes.CDLL(path)
    vp, cp, i, f, d = ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_float, ctypes.c_double
    u64 = ctypes.c_uint64
    sig = {
        "pd_create": (i, [cp, cp, cp, i, i, ctypes.POINTER(vp)]),
        "pd_create_synthetic": (i, [cp, cp, i, f, i, i, ctypes.POINTER(vp)]),
        "pd_destroy": (None, [vp]),
        "pd_compute_fat_points": (i, [cp, cp, i, vp, i]),
        "pd_last_error": (cp, [vp]),
        "pd_num_envs": (i, [vp]),
        "pd_state_words": (i, []),
        "pd_obs_dim": (i, []),
        "pd_car_state_bytes": (i, []),
        "pd_set_assists": (i, [vp, i, i, i]),
        "pd_set_tune": (i, [vp, cp, f]),
        "pd_set_raw_tune": (i, [vp, cp, f]),
        "pd_set_env_config": (i, [vp, vp]),
        "pd_get_env_config": (i, [vp, vp]),
        "pd_env_reset_counters": (i, [vp, vp]),
        "pd_params_bytes": (i, []),
        "pd_set_collision_response": (i, [vp, i]),
        "pd_get_contacts": (i, [vp, i, vp, i]),
        "pd_set_stream": (i, [vp, vp]),
        "pd_set_scoring_var": (i, [vp, cp, f]),
        "pd_get_scoring_var": (f, [vp, cp]),
        "pd_set_controls": (i, [vp, vp, vp, i, i]),
        "pd_set_actions": (i, [vp, vp, i]),
        "pd_step": (i, [vp, f, i]),
        "pd_get_time": (d, [vp]),
        "pd_set_time": (i, [vp, d]),
        "pd_teleport_spline": (i, [vp, vp, vp]),
        "pd_teleport_mode": (i, [vp, vp, i]),
        "pd_set_seed": (i, [vp, u64, u64]),
        "pd_get_car_state": (i, [vp, i, vp]),
        "pd_get_obs": (i, [vp, vp, i]),
        "pd_obs_dlpack": (vp, [vp]),
        "pd_observe": (i, [vp]),
        "pd_obs_device_ptr": (vp, [vp]),
        "pd_get_rewards": (i, [vp, vp, vp, vp]),
        "pd_env_step": (i, [vp, vp, f, vp, vp, vp]),
        "pd_env_step_host": (i, [vp, vp, f, vp, vp, vp]),
        "pd_tick_kernel": (ctypes.c_char_p, [vp]),
        "pd_tick_kernel_instance": (ctypes.c_char_p, [vp]),
        "pd_topology": (ctypes.c_int, [vp]),
        "pd_bvh_info": (i, [vp, vp, vp, vp]),
        "pd_env_stats": (i, [vp, vp, i]),
        "pd_set_autoreset": (i, [vp, i]),
        "pd_debug_read_clocks": (i, [vp, vp, i]),
        "pd_get_state": (i, [vp, i, vp]),
        "pd_set_state": (i, [vp, i, vp]),
        "pd_snapshot": (i, [vp, vp]),
        "pd_restore": (i, [vp, vp]),
        "pd_get_params": (i, [vp, vp]),
        "pd_get_track_info": (i, [vp, vp]),
        "pd_raycast": (i, [vp, i, vp, vp]),
        "pd_sync": (i, [vp]),
        "pd_stream": (vp, [vp]),
        "pd_launch_count": (u64, [vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _LIB = L
    return L


STATE_WORDS = None


def compute_fat_points(base_path, track, device=0, max_points=200000):
    """Track::computeFatPoints on the GPU: the content of spline.cache, [n, 15] float32 (best, left, right, center, forwardDir)."""
    L = load_library()
    out = np.zeros((max_points, 15), np.float32)
    n = L.pd_compute_fat_points(os.fspath(base_path).encode(), track.encode(), int(device), out.ctypes.data, max_points)
    if n < 0:
        msg = L.pd_last_error(None)
        raise PdError("pd_compute_fat_points failed (%d): %s" % (n, msg.decode() if msg else "?"))
    return out[:n].copy()


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    if hasattr(a, "data_ptr"):  # torch tensor
        return a.data_ptr()
    return a


class EnvConfig(ctypes.Structure):
    """PdEnvConfig of include/pd_batch.h: the ProjectDEnv knobs the kernels apply (projectd_env.py:27-53)."""
    _fields_ = [("min_gas", ctypes.c_float), ("max_gas", ctypes.c_float),
                ("terminate_hit_penalty", ctypes.c_float), ("terminate_off_track_penalty", ctypes.c_float), ("terminate_stuck_penalty", ctypes.c_float),
                ("terminate_low_reward", ctypes.c_float), ("stuck_timeout", ctypes.c_float),
                ("terminate_on_hit", ctypes.c_int32), ("terminate_off_track", ctypes.c_int32), ("terminate_when_stuck", ctypes.c_int32),
                ("smooth_controls", ctypes.c_int32), ("clutch", ctypes.c_float), ("requested_gear", ctypes.c_int32)]


class Batch:
    """N reference simulators with one car each, advanced together on one GPU.

    Mirrors what pyprojectd/projectd_env.py:118-136 does per environment: createSimulator + loadTrack + addCar.
    """

    def __init__(self, base_path, track="driftplayground", car="ks_toyota_ae86_drift", n_envs=1, device=0,
                 synthetic_tris=0, synthetic_length=20800.0):
        global STATE_WORDS
        self.L = load_library()
        STATE_WORDS = self.L.pd_state_words()
        self.words = STATE_WORDS
        self.h = ctypes.c_void_p()
        if synthetic_tris:
            rc = self.L.pd_create_synthetic(base_path.encode(), car.encode(), int(synthetic_tris), float(synthetic_length), int(n_envs), int(device), ctypes.byref(self.h))
        else:
            rc = self.L.pd_create(base_path.encode(), track.encode(), car.encode(), int(n_envs), int(device), ctypes.byref(self.h))
        if rc != 0:
            msg = self.L.pd_last_error(None)
            self.h = None
            raise PdError("pd_create failed (%d): %s" % (rc, msg.decode() if msg else "?"))
        self.n = n_envs
        self.device = device
        self._obs_tensor = None
        self._ctl = np.zeros((n_envs, 5), np.float32)
        self._gears = np.zeros((n_envs, 3), np.int8); self._gears[:, 0] = -1

    def controls_host(self):
        """Host mirror of the last controls handed over from host memory ([N,5] f32, [N,3] i8); the scalar
        setCarControls of the PyProjectD mirror edits one row of it."""
        return self._ctl, self._gears

    # -- lifetime -------------------------------------------------------------------------------------
    def close(self):
        self._obs_tensor = None          # the DLPack alias must not outlive the buffer it points at
        if getattr(self, "h", None):
            self.L.pd_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            msg = self.L.pd_last_error(self.h)
            raise PdError("libpd_b200 error %d: %s" % (rc, msg.decode() if msg else "?"))

    # -- configuration (PyProjectD.cpp:286-363) -------------------------------------------------------
    def set_assists(self, auto_clutch=True, auto_shift=True, auto_blip=True):
        self._ck(self.L.pd_set_assists(self.h, int(auto_clutch), int(auto_shift), int(auto_blip)))

    def set_tune(self, name, value):
        self._ck(self.L.pd_set_tune(self.h, name.encode(), float(value)))

    def set_raw_tune(self, name, value):
        self._ck(self.L.pd_set_raw_tune(self.h, name.encode(), float(value)))

    def env_config(self):
        c = EnvConfig()
        self._ck(self.L.pd_get_env_config(self.h, ctypes.byref(c)))
        return c

    def set_env_config(self, **kw):
        """Update fields of the batch's PdEnvConfig (names as in ProjectDEnv: min_gas, terminate_hit_penalty, ...)."""
        c = self.env_config()
        for k, v in kw.items():
            if not hasattr(c, k):
                raise TypeError("unknown env config field %r" % k)
            setattr(c, k, v)
        self._ck(self.L.pd_set_env_config(self.h, ctypes.byref(c)))
        return c

    def env_reset_counters(self, mask=None):
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        self._ck(self.L.pd_env_reset_counters(self.h, _ptr(m)))

    def set_collision_response(self, on=True):
        """Contact joints from collisions enter the solve (default) / collisions only raise collisionFlag."""
        self._ck(self.L.pd_set_collision_response(self.h, int(bool(on))))

    def contacts(self, env=0):
        """Live contact joints of one env: array [k, 8] = position, normal, depth, kind."""
        out = np.zeros((8, 8), np.float32)
        k = self.L.pd_get_contacts(self.h, int(env), out.ctypes.data, 8)
        if k < 0:
            raise PdError("pd_get_contacts failed")
        return out[:k]

    def set_stream(self, stream_ptr):
        """Run this batch's kernels on the given cudaStream_t (int / None = the batch's own stream)."""
        self._ck(self.L.pd_set_stream(self.h, ctypes.c_void_p(stream_ptr) if stream_ptr else None))

    def set_scoring_var(self, name, value):
        self._ck(self.L.pd_set_scoring_var(self.h, name.encode(), float(value)))

    def get_scoring_var(self, name):
        return self.L.pd_get_scoring_var(self.h, name.encode())

    def set_seed(self, seed, env_id_offset=0):
        self._ck(self.L.pd_set_seed(self.h, int(seed), int(env_id_offset)))

    # -- controls / stepping --------------------------------------------------------------------------
    def set_controls(self, controls, gears=None, smooth=True):
        """controls [N,5] float32 (steer, clutch, brake, handBrake, gas); gears [N,3] int8 or None."""
        on_dev = hasattr(controls, "data_ptr")
        if not on_dev:
            controls = np.ascontiguousarray(controls, dtype=np.float32)
            assert controls.shape == (self.n, 5)
            if gears is not None:
                gears = np.ascontiguousarray(gears, dtype=np.int8)
                if gears is not self._gears:
                    self._gears[...] = gears
            if controls is not self._ctl:
                self._ctl[...] = controls
        self._ck(self.L.pd_set_controls(self.h, _ptr(controls), _ptr(gears), int(smooth), int(on_dev)))

    def set_actions(self, actions):
        on_dev = hasattr(actions, "data_ptr")
        if not on_dev:
            actions = np.ascontiguousarray(actions, dtype=np.float32)
            assert actions.shape == (self.n, 2)
        self._ck(self.L.pd_set_actions(self.h, _ptr(actions), int(on_dev)))

    def step(self, dt=1.0 / 333.0, n_ticks=1):
        self._ck(self.L.pd_step(self.h, float(dt), int(n_ticks)))

    def env_step(self, actions_dev, dt=1.0 / 333.0, obs=None, reward=None, done=None):
        self._ck(self.L.pd_env_step(self.h, _ptr(actions_dev), float(dt), _ptr(obs), _ptr(reward), _ptr(done)))

    def env_step_host(self, actions, dt=1.0 / 333.0, obs=None, reward=None, done=None):
        """One env step for host-resident buffers (numpy arrays or CPU torch tensors, ideally pinned):
        H2D actions, step, D2H obs / reward / done, one stream sync."""
        self._ck(self.L.pd_env_step_host(self.h, _ptr(actions), float(dt), _ptr(obs), _ptr(reward), _ptr(done)))

    def set_autoreset(self, mode):
        """0 = same step (3 launches per env step), 1 = next step (1 launch; gymnasium >= 1.0 VectorEnv convention)."""
        self._ck(self.L.pd_set_autoreset(self.h, int(mode)))

    def debug_warp_clocks(self):
        out = np.zeros(4096 + self.n * 16, dtype=np.int64)
        cnt = self.L.pd_debug_read_clocks(self.h, out.ctypes.data, out.size)
        return out[:cnt]

    def env_stats(self, reset=True):
        out = np.zeros(8, dtype=np.float64)
        self._ck(self.L.pd_env_stats(self.h, out.ctypes.data, int(reset)))
        return out

    def time(self):
        return self.L.pd_get_time(self.h)

    def set_time(self, t):
        self._ck(self.L.pd_set_time(self.h, float(t)))

    def sync(self):
        self._ck(self.L.pd_sync(self.h))

    def tick_kernel(self):
        return self.L.pd_tick_kernel(self.h).decode()

    def tick_kernel_instance(self):
        return self.L.pd_tick_kernel_instance(self.h).decode()

    def bvh_info(self):
        """(built on the device?, number of nodes, depth of the device-built tree)"""
        a, b_, c = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_int(0)
        self._ck(self.L.pd_bvh_info(self.h, ctypes.byref(a), ctypes.byref(b_), ctypes.byref(c)))
        return bool(a.value), b_.value, c.value

    def topology(self):
        """(front == DWB) * 2 + (rear == DWB)"""
        return int(self.L.pd_topology(self.h))

    def launch_count(self):
        return int(self.L.pd_launch_count(self.h))

    def stream(self):
        return self.L.pd_stream(self.h)

    # -- teleports ------------------------------------------------------------------------------------
    def teleport_spline(self, dist_norm=None, mask=None):
        d = None if dist_norm is None else np.ascontiguousarray(np.broadcast_to(np.asarray(dist_norm, dtype=np.float32), (self.n,)), dtype=np.float32)
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        self._ck(self.L.pd_teleport_spline(self.h, _ptr(m), _ptr(d)))

    def teleport_mode(self, mode=0, mask=None):
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        self._ck(self.L.pd_teleport_mode(self.h, _ptr(m), int(mode)))

    # -- outputs --------------------------------------------------------------------------------------
    def obs_host(self):
        out = np.zeros((self.n, OBS_DIM), dtype=np.float32)
        self._ck(self.L.pd_get_obs(self.h, out.ctypes.data, 0))
        return out

    def obs_dlpack_capsule(self):
        ptr = self.L.pd_obs_dlpack(self.h)
        if not ptr:
            raise PdError("pd_obs_dlpack failed")
        new = ctypes.pythonapi.PyCapsule_New
        new.restype = ctypes.py_object
        new.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p]
        return new(ptr, b"dltensor", None)

    def obs_tensor(self):
        """torch CUDA tensor [N,24] float32 aliasing the batch's observation buffer (DLPack hand-off)."""
        import torch
        if self._obs_tensor is None:
            self._obs_tensor = torch.from_dlpack(self.obs_dlpack_capsule())
        return self._obs_tensor

    def observe(self):
        self._ck(self.L.pd_observe(self.h))

    def rewards(self):
        r = np.zeros(self.n, np.float32); t = np.zeros(self.n, np.float32); fl = np.zeros(self.n, np.int32)
        self._ck(self.L.pd_get_rewards(self.h, r.ctypes.data, t.ctypes.data, fl.ctypes.data))
        return r, t, fl

    def car_state_bytes(self, env=0):
        buf = np.zeros(self.L.pd_car_state_bytes(), dtype=np.uint8)
        self._ck(self.L.pd_get_car_state(self.h, int(env), buf.ctypes.data))
        return buf

    # -- state records --------------------------------------------------------------------------------
    def get_state(self, env=0):
        rec = np.zeros(self.words, dtype=np.uint32)
        self._ck(self.L.pd_get_state(self.h, int(env), rec.ctypes.data))
        return rec

    def set_state(self, env, rec):
        rec = np.ascontiguousarray(rec, dtype=np.uint32)
        assert rec.shape == (self.words,)
        self._ck(self.L.pd_set_state(self.h, int(env), rec.ctypes.data))

    def snapshot(self):
        """All states as [words, N] uint32 (the device's structure-of-arrays layout)."""
        buf = np.zeros((self.words, self.n), dtype=np.uint32)
        self._ck(self.L.pd_snapshot(self.h, buf.ctypes.data))
        return buf

    def restore(self, buf):
        buf = np.ascontiguousarray(buf, dtype=np.uint32)
        assert buf.shape == (self.words, self.n)
        self._ck(self.L.pd_restore(self.h, buf.ctypes.data))

    def params_bytes(self):
        buf = np.zeros(self.L.pd_params_bytes(), dtype=np.uint8)
        self._ck(self.L.pd_get_params(self.h, buf.ctypes.data))
        return buf

    def track_info(self):
        buf = np.zeros(12, dtype=np.uint32)
        self._ck(self.L.pd_get_track_info(self.h, buf.ctypes.data))
        i = buf[:8].view(np.int32); f = buf[8:].view(np.float32)
        return dict(nSurfaces=int(i[0]), nTris=int(i[1]), nNodes=int(i[2]), nFatPoints=int(i[3]), nSplineNodes=int(i[4]),
                    interpolateStep=int(i[5]), closedLoop=int(i[6]), computedTrackLength=float(f[0]), computedTrackWidth=float(f[1]),
                    dynamicGripLevel=float(f[2]), hashCellSize=float(f[3]))

    def raycast(self, rays):
        rays = np.ascontiguousarray(rays, dtype=np.float32)
        out = np.zeros((rays.shape[0], 8), dtype=np.float32)
        self._ck(self.L.pd_raycast(self.h, rays.shape[0], rays.ctypes.data, out.ctypes.data))
        return out
