"""Prints every field that leaves the single-tick tolerance on yamanashi_short's off-track drives (GPU vs oracle)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import pdref
from parity_util import *
from projectd_core_b200 import Batch
lay = pdref.Layout(); track = "yamanashi_short"; n = 64
b = make_env_like(Batch(pdref.BASE_PATH, track=track, n_envs=n, device=0))
starts = drive_start_states(pdref, lay, track, n)
refs = [pdref.RefSim(track=track) for _ in range(n)]
for r, (rec, tm, fr) in zip(refs, starts):
    r.set_state(rec); r.set_time(0.0)
recs = [r.state() for r in refs]
shown = 0
for t in range(450):
    for i, r in enumerate(refs):
        r.set_controls(**drive_controls(t, i, lay, recs[i]))
    before = [r.state().copy() for r in refs]
    b.restore(np.stack(before, axis=1)); b.set_time(refs[0].time()); b.step(1 / 333.0, 1)
    out = b.snapshot()
    for i, r in enumerate(refs):
        r.step(); recs[i] = r.state()
        bad, w = compare_records(lay, out[:, i], recs[i], tol=1e-4)
        if bad and shown < 12:
            shown += 1
            print("tick", t, "env", i, "worst", w)
            for x in bad[:16]:
                print("   ", x)
            for wq in range(4):
                for fld in ("contactX", "contactY", "contactZ", "depth", "load", "normalX", "normalY", "normalZ", "distToGround", "loadedRadius"):
                    a = lay.get(out[:, i], "tyre%d.%s" % (wq, fld)); bq = lay.get(recs[i], "tyre%d.%s" % (wq, fld))
                    if a != bq:
                        print("      tyre%d.%s mine %.9g ref %.9g diff %.3g" % (wq, fld, a, bq, a - bq))
            for wq in range(4):
                print("    tyre%d surf %d contact %d load %.3f depth %.6f ndSlip %.4f dirty %.5f" % (wq, lay.get(recs[i], "tyre%d.surfaceId" % wq), lay.get(recs[i], "tyre%d.hasContact" % wq), lay.get(recs[i], "tyre%d.load" % wq), lay.get(recs[i], "tyre%d.depth" % wq), lay.get(recs[i], "tyre%d.ndSlip" % wq), lay.get(recs[i], "tyre%d.dirtyLevel" % wq)))
