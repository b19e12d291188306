"""Debug aid: the single-tick parity loop of tests/test_gpu_parity.py for one car, printing every field of the first records
that leave the rule.  python tools/debug_car_parity.py CAR [N_ENVS] [TICKS]"""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pdref as oracle
from projectd_core_b200 import Batch
from parity_util import arbitrate, compare_records, drive_controls, drive_start_states, make_env_like
car = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 48; ticks = int(sys.argv[3]) if len(sys.argv) > 3 else 400
lay = oracle.Layout(); DT = 1.0 / 333.0
b = make_env_like(Batch(oracle.BASE_PATH, n_envs=n, device=0, car=car))
print(b.tick_kernel_instance())
starts = drive_start_states(oracle, lay, "driftplayground", n, car=car)
refs = [oracle.RefSim(car=car) for _ in range(n)]
for r, (rec, tm, fr) in zip(refs, starts):
    r.set_state(rec); r.set_time(0.0)
recs = [r.state() for r in refs]; shown = 0
for t in range(ticks):
    for i, r in enumerate(refs):
        r.set_controls(**drive_controls(t, i, lay, recs[i]))
    before = [r.state() for r in refs]; tb = refs[0].time()
    b.restore(np.stack(before, axis=1)); b.set_time(tb); b.step(DT, 1)
    out = b.snapshot()
    for i, r in enumerate(refs):
        ncb = len(r.contacts())
        r.step(); ref = recs[i] = r.state()
        bad, w = compare_records(lay, out[:, i], ref, tol=1e-4)
        if bad:
            left = arbitrate(oracle, lay, "driftplayground", before[i], tb, ref, bad, car=car)
            if left and shown < 4:
                shown += 1
                print("tick", t, "env", i, "frame before", lay.get(before[i], "car.physFrame"), "oracle contacts before/after", ncb, len(r.contacts()),
                      "collisionFlag mine/ref", lay.get(out[:, i], "car.collisionFlag"), lay.get(ref, "car.collisionFlag"), "speed", lay.get(ref, "car.speed"))
                for x in left: print("   ", x)
                print("    oracle contacts:", r.contacts())
print("done")
