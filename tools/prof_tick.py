"""Small driver for ncu: N envs, warm-up, then a few plain ticks with random actions (no oracle, no torch needed)."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pdref
from projectd_core_b200 import Batch
from projectd_core_b200.env import configure_like_env as make_env_like      # the bench configuration: env that terminates on hit -> detection only, collision warp inside the tick kernel
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
ticks = int(sys.argv[2]) if len(sys.argv) > 2 else 40
b = make_env_like(Batch(pdref.BASE_PATH, n_envs=n, device=0)); b.set_seed(1234, 0); b.teleport_mode(2)
rng = np.random.default_rng(0)
for t in range(ticks):
    if t % 33 == 0:
        b.set_actions(rng.uniform(-1, 1, (n, 2)).astype(np.float32))
    b.step(1.0 / 333.0, 1)
b.sync()
print("done", b.launch_count())
