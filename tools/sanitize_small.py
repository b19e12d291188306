"""Small run for compute-sanitizer: both tick kernels, both auto-reset conventions, collisions, teleports."""
import sys, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pdref
from projectd_core_b200 import Batch
from parity_util import make_env_like
for quad_max, n in ((8192, 37), (0, 70)):
    os.environ["PD_QUAD_MAX_ENVS"] = str(quad_max)
    for mode in (0, 1):
        b = make_env_like(Batch(pdref.BASE_PATH, n_envs=n, device=0)); b.set_seed(3, 0); b.teleport_mode(2); b.set_autoreset(mode)
        act = torch.zeros((n, 2), device="cuda"); act[:, 0] = torch.linspace(-1, 1, n, device="cuda"); act[:, 1] = 1.0
        rew = torch.zeros(n, device="cuda"); done = torch.zeros(n, device="cuda", dtype=torch.int32)
        for t in range(40):
            b.env_step(act, 1.0 / 333.0, None, rew, done)
        b.step(1.0 / 333.0, 3); b.observe(); b.sync()
        print(b.tick_kernel(), "mode", mode, "ok", float(rew.sum()), int((done != 0).sum())); b.close()
# round 2: the other cars (double-wishbone kernel instances, turbo), both kernel families, detection-only and response collisions
from projectd_core_b200.env import configure_like_env
for car in ("ks_mazda_rx7_tuned", "dthwsh_mazda_rx7_fc3s_sr20"):
    for quad_max, n in ((20480, 21), (0, 45)):
        os.environ["PD_QUAD_MAX_ENVS"] = str(quad_max)
        for cfg in (make_env_like, configure_like_env):
            b = cfg(Batch(pdref.BASE_PATH, n_envs=n, device=0, car=car)); b.set_seed(5, 0); b.teleport_mode(2); b.set_autoreset(1)
            act = torch.zeros((n, 2), device="cuda"); act[:, 0] = torch.linspace(-1, 1, n, device="cuda"); act[:, 1] = 1.0
            rew = torch.zeros(n, device="cuda"); done = torch.zeros(n, device="cuda", dtype=torch.int32)
            for t in range(30):
                b.env_step(act, 1.0 / 333.0, None, rew, done)
            b.sync()
            print(car, b.tick_kernel_instance(), "ok", float(rew.sum())); b.close()

