"""BASELINE configs[0] on the GPU: the bundled demo car, 1 env, scripted throttle / steer, 10 000 ticks, free running;
position / heading / speed difference to the oracle's stored trajectory (tests/golden/demo_golden.npz, a record every 250 ticks)."""
import sys, os, math
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pdref
from projectd_core_b200 import Batch
from parity_util import make_env_like
g = np.load(os.path.join(ROOT, "tests", "golden", "demo_golden.npz")); lay = pdref.Layout()
b = make_env_like(Batch(pdref.BASE_PATH, n_envs=1, device=0)); b.teleport_spline(0.0)
ctl = np.zeros((1, 5), np.float32)
print("tick   |dpos| m   dyaw deg   speed gpu / ref m/s   gear gpu/ref   collisionFlag gpu/ref")
for t in range(10001):
    if t % 250 == 0:
        rec = b.get_state(0); ref = g["traj_state"][t // 250]
        dp = math.sqrt(sum((lay.get(rec, "chassis.p" + k) - lay.get(ref, "chassis.p" + k)) ** 2 for k in "xyz"))
        yaw = lambda r: math.atan2(lay.get(r, "chassis.azx"), lay.get(r, "chassis.azz"))
        dy = math.degrees((yaw(rec) - yaw(ref) + math.pi) % (2 * math.pi) - math.pi)
        sp = lambda r: math.sqrt(sum(lay.get(r, "chassis.v" + k) ** 2 for k in "xyz"))
        if t in (0, 250, 500, 750, 1000, 1500, 2000, 3000, 5000, 7500, 10000):
            print("%5d  %9.4f  %9.3f   %7.2f / %7.2f      %d / %d          %d / %d" % (t, dp, dy, sp(rec), sp(ref), lay.get(rec, "car.currentGear"), lay.get(ref, "car.currentGear"), lay.get(rec, "car.collisionFlag"), lay.get(ref, "car.collisionFlag")))
    if t == 10000: break
    ctl[0, 0] = 0.3 * math.sin(2 * math.pi * t / 999.0); ctl[0, 4] = 0.1 + 0.9 * min(1.0, t / 333.0)
    b.set_controls(ctl, None, True); b.step(1.0 / 333.0, 1)
