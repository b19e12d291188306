"""Per-warp duration spread of the tick kernel in the rollout's steady state (needs PD_DEBUG_CLOCKS=1).
python tools/warp_tail.py ENVS PREROLL"""
import sys, os
os.environ["PD_DEBUG_CLOCKS"] = "1"
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pdref
from projectd_core_b200 import Batch
from projectd_core_b200.env import configure_like_env as make_env_like      # the bench configuration: env that terminates on hit -> detection only, collision warp inside the tick kernel
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
pre = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
dev = torch.device("cuda", 0)
b = make_env_like(Batch(pdref.BASE_PATH, n_envs=n, device=0)); b.set_seed(1234, 0); b.teleport_mode(2); b.set_autoreset(1)
lay = pdref.Layout()
gen = torch.Generator(device=dev); gen.manual_seed(5)
rew = torch.zeros(n, device=dev); done = torch.zeros(n, device=dev, dtype=torch.int32)
for t in range(pre + 8):
    if t % 33 == 0:
        a = (torch.rand((n, 2), device=dev, generator=gen) * 2 - 1).contiguous(); torch.cuda.synchronize()
    if t >= pre:
        before = b.snapshot()
    b.env_step(a, 1.0 / 333.0, None, rew, done)
    if t >= pre or t in (5, 100, 500, 1000):
        c = b.debug_warp_clocks().astype(np.float64)
        c = c[c > 0]
        q = np.percentile(c, [50, 90, 99, 100])
        print("tick %d: warps %d  p50 %.0f  p90 %.0f  p99 %.0f  max %.0f kcycles (max/p50 %.2f)  done %d" % (t, len(c), q[0] / 1e3, q[1] / 1e3, q[2] / 1e3, q[3] / 1e3, q[3] / q[0], int((done != 0).sum())))
        if t >= pre:
            cars_per_warp = n / len(c)
            for w in np.argsort(-c)[:3]:
                e0 = int(w * cars_per_warp); e1 = int((w + 1) * cars_per_warp)
                info = []
                for e in range(e0, min(e1, n)):
                    r = before[:, e]
                    v = np.array([lay.get(r, "chassis.v" + k) for k in "xyz"])
                    info.append("e%d v=%.1f gear=%d oot=%d pend?" % (e, float(np.linalg.norm(v)), lay.get(r, "car.currentGear"), lay.get(r, "car.outOfTrackFlag")))
                print("   slow warp %d (%.0f kcycles): %s" % (w, c[w] / 1e3, "; ".join(info)))
