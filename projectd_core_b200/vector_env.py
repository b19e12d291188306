"""gymnasium.vector.VectorEnv over the batched core -- the vectorised counterpart of the reference's gymnasium wrapper
(pyprojectd/projectd_gymnasium/projectd_gymnasium.py:9-36, id "ProjectD-v0", max_episode_steps 80000).

* With gymnasium installed the class IS a ``gymnasium.vector.VectorEnv`` (spaces are ``gymnasium.spaces.Box``); without it the
  same class stands on a minimal ``Box`` with the attributes vector-env consumers read (low / high / shape / dtype / sample).
* Auto-reset follows gymnasium >= 1.0's NEXT_STEP mode: the step that ends an episode returns its terminal observation, the
  next step ignores that env's action and returns the reset observation -- which is exactly PD_AUTORESET_NEXT_STEP, i.e. ONE
  kernel launch per vector step.
* observations / rewards / terminated / truncated are torch CUDA tensors (the observation aliases the library's buffer through
  DLPack; pass ``to_numpy=True`` for host arrays).  ``truncated`` fires at ``max_episode_steps`` (80000 like the registration).
"""
from __future__ import annotations

import numpy as np

from .env import BatchedProjectDEnv

try:                                     # pragma: no cover - gymnasium is absent from the build image
    import gymnasium as _gym
    from gymnasium.spaces import Box
    from gymnasium.vector import VectorEnv as _Base
    from gymnasium.vector.utils import batch_space as _batch_space
    HAVE_GYMNASIUM = True
except Exception:                        # minimal stand-ins with the attributes consumers read
    HAVE_GYMNASIUM = False

    class Box:
        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.low = np.asarray(low, dtype=dtype); self.high = np.asarray(high, dtype=dtype)
            self.shape = tuple(self.low.shape if shape is None else shape); self.dtype = np.dtype(dtype)
            self._rng = np.random.default_rng()

        def seed(self, seed=None):
            self._rng = np.random.default_rng(seed)

        def sample(self):
            return self._rng.uniform(self.low, self.high).astype(self.dtype)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

    class _Base:
        metadata = {"autoreset_mode": "NextStep"}

    def _batch_space(space, n):
        return Box(np.broadcast_to(space.low, (n,) + space.shape).copy(), np.broadcast_to(space.high, (n,) + space.shape).copy(), dtype=space.dtype)


class ProjectDVectorEnv(_Base):
    """N ``ProjectD-v0`` environments on one GPU behind the VectorEnv interface."""

    metadata = {"render_modes": [], "autoreset_mode": "NextStep"}
    max_episode_steps = 80000            # projectd_gymnasium/__init__.py:7

    def __init__(self, base_dir, num_envs=1024, device=0, seed=0, to_numpy=False, **env_kwargs):
        env_kwargs.setdefault("autoreset_mode", 1)
        if env_kwargs["autoreset_mode"] != 1:
            raise ValueError("the VectorEnv interface uses the next-step auto-reset (gymnasium >= 1.0)")
        self.impl = BatchedProjectDEnv(base_dir, num_envs=num_envs, device=device, seed=seed, **env_kwargs)
        self.num_envs = int(num_envs)
        self.to_numpy = bool(to_numpy)
        lo, hi = self.impl.observation_bounds(); alo, ahi = self.impl.action_bounds()
        self.single_observation_space = Box(lo, hi, dtype=np.float32)
        self.single_action_space = Box(alo, ahi, dtype=np.float32)
        self.observation_space = _batch_space(self.single_observation_space, self.num_envs)
        self.action_space = _batch_space(self.single_action_space, self.num_envs)
        import torch
        self._steps = torch.zeros(self.num_envs, dtype=torch.int32, device=self.impl.device)
        self._was_done = torch.zeros(self.num_envs, dtype=torch.bool, device=self.impl.device)

    def _out(self, t):
        return t.cpu().numpy() if self.to_numpy else t

    def reset(self, *, seed=None, options=None):
        if seed is not None:
            self.impl.batch.set_seed(int(seed), 0)
        obs = self.impl.reset()
        self._steps.zero_(); self._was_done.zero_()
        return self._out(obs), {}

    def step(self, actions):
        import torch
        obs, rew, term, _, _ = self.impl.step(actions)
        # envs that finished at the previous step were reset by this step (their action was ignored): their step counter restarts
        self._steps = torch.where(self._was_done, torch.zeros_like(self._steps), self._steps + 1)
        trunc = (self._steps >= self.max_episode_steps) & ~term
        self._was_done = term | trunc
        if bool(trunc.any()):           # rare (80000 steps): a truncated env is reset at the next step like a terminated one
            self.impl.batch.teleport_mode(self.impl.teleport_mode, trunc.to(torch.uint8).cpu().numpy())
        return self._out(obs), self._out(rew), self._out(term), self._out(trunc), {}

    def close(self, **kwargs):
        self.impl.close()

    # gymnasium.vector.VectorEnv API completeness
    def close_extras(self, **kwargs):
        self.impl.close()

    @property
    def unwrapped(self):
        return self


def make_vec(base_dir, num_envs=1024, **kwargs):
    """``gymnasium.make_vec("ProjectD-v0", num_envs=...)`` for the batched core."""
    return ProjectDVectorEnv(base_dir, num_envs=num_envs, **kwargs)
