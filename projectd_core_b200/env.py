"""BatchedProjectDEnv -- the reference's ``ProjectDEnv`` (pyprojectd/projectd_env.py) vectorised over an env batch.

Same configuration surface (class attributes with the reference's names and defaults), same observation (24 floats,
projectd_env.py:237-275), same action mapping (steer = a0, gas = linscale(a1, -1..1 -> min_gas..max_gas), :158-160),
same reward / termination rule (:178-212: stepReward minus the hit / off-track / stuck penalties, low-reward cut).
What changes is the shape: N environments advance per call on one GPU, observations / rewards / done flags are
CUDA tensors (DLPack hand-off of the library's own buffers) and terminated envs are reset inside the step
(teleport + one zero-action tick, exactly what ``reset()`` does in the reference, :218-230).
"""
from __future__ import annotations

import math
import os

import numpy as np

from .binding import Batch, OBS_DIM
from . import dist as pdist


def configure_like_env(batch, car_model="ks_toyota_ae86_drift", auto_clutch=True, auto_shift=True, auto_blip=True):
    """Configure a Batch the way ProjectDEnv.__init__ configures its simulator (projectd_env.py:118-136): assists, the
    car's tune table, the scoring variables."""
    batch.set_assists(auto_clutch, auto_shift, auto_blip)
    batch.set_collision_response(not BatchedProjectDEnv.terminate_on_hit)        # as BatchedProjectDEnv does for an env that terminates on hit
    for name, value in BatchedProjectDEnv.car_tunes.get(car_model, {}).items():
        batch.set_tune(name, value)
    for name, value in BatchedProjectDEnv.scoring_vars.items():
        batch.set_scoring_var(name, value)
    return batch


class BatchedProjectDEnv:
    sim_dt = 1.0 / 333.0
    track_name = "driftplayground"
    car_model = "ks_toyota_ae86_drift"

    smooth_controls = True
    auto_clutch = True
    auto_shift = True
    auto_blip = True

    range_velocity = 100
    range_angularVelocity = 100
    range_tyreNdSlip = 10
    range_lookAhead = math.pi
    range_probe = 50

    terminate_on_hit = True
    terminate_off_track = True
    terminate_when_stuck = True

    terminate_hit_penalty = 50.0
    terminate_off_track_penalty = 50.0
    terminate_stuck_penalty = 50.0
    terminate_low_reward = -200.0
    stuck_timeout = 5.0

    collision_response = None   # None: on unless terminate_on_hit (the terminal tick then shows the flag but not the bounce); True / False: forced
    teleport_mode = 0   # 0:Start, 1:Nearest, 2:Random
    autoreset_mode = 0  # 0: a finished env is reset inside the step that finished it (SB3-style VecEnv, 3 launches per step);
                        # 1: at the following step, whose action is ignored (gymnasium >= 1.0 VectorEnv; 1 launch per step)
    min_gas = 0.1
    max_gas = 1.0

    scoring_vars = {
        "SmoothSteerSpeed": 10.0, "MinBonusSpeed": 5.0, "MaxBonusSpeed": 200.0, "StallRpm": 300.0,
        "DirectionThreshold": 0.75, "OutOfTrackThreshold": 0.51, "ApproachDistance": 3.5, "CriticalDistance": 2.0,
        "TravelBonus": 0.1, "TravelSplineBonus": 0.01, "DriftBonus": 0.0, "SpeedBonus": 0.0, "ThrottleBonus": 0.0,
        "EngineRpmBonus": 0.0, "DirectionBonus": 0.0, "DirectionPenalty": 0.0, "ObstApproachPenalty": 0.0,
        "CollisionPenalty": 0.0, "OffTrackPenalty": 0.0, "GearGrindPenalty": 0.0, "StallPenalty": 0.0,
    }
    car_tunes = {
        "ks_toyota_ae86_drift": {"FRONT_BIAS": 55.0, "DIFF_POWER": 30.0, "DIFF_COAST": 30.0, "FINAL_RATIO": 5.0,
                                 "PRESSURE_LF": 28.0, "PRESSURE_RF": 28.0, "PRESSURE_LR": 28.0, "PRESSURE_RR": 28.0},
    }

    def __init__(self, base_dir, num_envs=1024, device=0, seed=0, total_envs=None, env_id_offset=0, **kwargs):
        """base_dir: directory holding cfg/ and content/ (the reference checkout).  For a sharded run pass this
        rank's slice (see projectd_core_b200.dist.shard_range) as num_envs / env_id_offset."""
        for k, v in kwargs.items():
            if not hasattr(type(self), k):
                raise TypeError("unknown option %r" % k)
            setattr(self, k, v)
        self.base_dir = os.fspath(base_dir)
        self.num_envs = int(num_envs)
        self.batch = Batch(self.base_dir, track=self.track_name, car=self.car_model, n_envs=self.num_envs, device=device)
        b = self.batch
        b.set_seed(seed, env_id_offset)
        b.teleport_mode(self.teleport_mode)                                    # projectd_env.py:123
        b.set_autoreset(self.autoreset_mode)
        b.set_assists(self.auto_clutch, self.auto_shift, self.auto_blip)        # :125
        for name, value in self.car_tunes.get(self.car_model, {}).items():      # :127-129
            b.set_tune(name, value)
        for name, value in self.scoring_vars.items():                           # :131-132
            b.set_scoring_var(name, value)
        # the env-level knobs live in the kernels (PdEnvConfig): gas range, penalties, termination switches, the clutch / gear
        # overrides of projectd_env.py:162-166 -- every attribute above is live, none is a dead knob
        b.set_env_config(min_gas=self.min_gas, max_gas=self.max_gas, terminate_hit_penalty=self.terminate_hit_penalty,
                         terminate_off_track_penalty=self.terminate_off_track_penalty, terminate_stuck_penalty=self.terminate_stuck_penalty,
                         terminate_low_reward=self.terminate_low_reward, stuck_timeout=self.stuck_timeout,
                         terminate_on_hit=int(self.terminate_on_hit), terminate_off_track=int(self.terminate_off_track),
                         terminate_when_stuck=int(self.terminate_when_stuck), smooth_controls=int(self.smooth_controls),
                         clutch=0.0 if self.auto_clutch else 1.0, requested_gear=-1 if self.auto_shift else 2)
        # collision response (contact joints): needed when an env lives on after a hit; an env that terminates on hit only needs
        # the flag, and the detection then rides inside the tick kernel (one launch per step)
        b.set_collision_response(self.collision_response if self.collision_response is not None else not self.terminate_on_hit)
        import torch
        self.device = torch.device("cuda", device)
        # stream ordering: the batch launches on its own non-blocking stream; every step makes that stream wait for the caller's
        # current stream (the producer of `actions`) and the caller's stream wait for the tick (before obs / reward / done are read)
        self._stream = torch.cuda.ExternalStream(b.stream(), device=self.device)
        self.obs = b.obs_tensor()                                               # [N,24] view of the library's buffer
        self.reward = torch.zeros(self.num_envs, device=self.device)
        self.done = torch.zeros(self.num_envs, device=self.device, dtype=torch.int32)
        self._zero_action = torch.zeros((self.num_envs, 2), device=self.device)
        self.step_id = 0

    # ---- spaces (projectd_env.py:279-360) ----
    def observation_bounds(self):
        lo = np.array([-self.range_velocity] * 3 + [-self.range_angularVelocity] * 3 + [0.0] * 4 + [-1.0, -1.0]
                      + [-self.range_lookAhead] * 5 + [0.0] * 7, dtype=np.float32)
        hi = np.array([self.range_velocity] * 3 + [self.range_angularVelocity] * 3 + [self.range_tyreNdSlip] * 4 + [1.0, 1.0]
                      + [self.range_lookAhead] * 5 + [self.range_probe] * 7, dtype=np.float32)
        assert lo.size == OBS_DIM
        return lo, hi

    def action_bounds(self):
        return np.array([-1.0, -1.0], np.float32), np.array([1.0, 1.0], np.float32)

    # ---- gym-style API, vectorised ----
    def reset(self):
        """Teleports every env by `teleport_mode` and advances one zero-action tick (projectd_env.py:218-230)."""
        import torch
        cur = torch.cuda.current_stream(self.device)
        self._stream.wait_stream(cur)
        self.batch.teleport_mode(self.teleport_mode)
        self.batch.env_step(self._zero_action, self.sim_dt, None, self.reward, self.done)
        self.batch.env_reset_counters()            # total_reward = 0 after the reset's own step (projectd_env.py:224-227)
        self.batch.env_stats(reset=True)
        cur.wait_stream(self._stream)
        self.step_id = 0
        return self.obs

    def step(self, actions):
        """actions: CUDA float32 tensor [N,2] in [-1,1] (steer, gas).  Returns (obs, reward, terminated, truncated, info);
        all tensors live on the GPU and alias buffers that the next step overwrites."""
        import torch
        if not (isinstance(actions, torch.Tensor) and actions.is_cuda):
            actions = torch.as_tensor(np.asarray(actions, dtype=np.float32), device=self.device)
        actions = actions.to(torch.float32).contiguous()
        if actions.shape != (self.num_envs, 2):
            raise ValueError("actions must be [num_envs, 2]")
        cur = torch.cuda.current_stream(self.device)
        self._stream.wait_stream(cur)              # the tick kernel reads `actions` only after the caller's stream has produced them
        self.batch.env_step(actions, self.sim_dt, None, self.reward, self.done)
        # obs / reward / done are complete before anything on the caller's stream reads them -- and the memory of `actions` is
        # not recycled under the kernel either: the caching allocator hands a freed block to later work of the caller's stream
        # only, all of which is ordered after this wait (no record_stream: it would make the allocator touch the batch's stream
        # after close() has destroyed it)
        cur.wait_stream(self._stream)
        self.step_id += 1
        return self.obs, self.reward, self.done.bool(), torch.zeros_like(self.done, dtype=torch.bool), {}

    def episode_stats(self, reset=True, all_ranks=True):
        """Episode statistics since the last call; summed over ranks when a process group is up (the path's only
        collective, once per rollout)."""
        s = self.batch.env_stats(reset=reset)
        if all_ranks:
            s = pdist.reduce_stats(s, self.device)
        return pdist.summarize(s)

    def close(self):
        self.obs = None                            # alias of the batch's observation buffer: gone with the batch
        self.batch.close()
