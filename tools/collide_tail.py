"""Which cars make k_collide slow?  (profiling build with per-car clocks: libpd_b200_dbg.so)"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ["PD_DEBUG_CLOCKS"] = "1"
os.environ["PD_B200_LIB"] = os.path.join(ROOT, "projectd_core_b200", "libpd_b200_dbg.so")
import numpy as np, torch
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pdref
from projectd_core_b200 import Batch
from projectd_core_b200.env import configure_like_env as make_env_like      # the bench configuration: env that terminates on hit -> detection only, collision warp inside the tick kernel
n = 4096; pre = 2001
dev = torch.device("cuda", 0)
b = make_env_like(Batch(pdref.BASE_PATH, n_envs=n, device=0)); b.set_seed(1234, 0); b.teleport_mode(2); b.set_autoreset(1)
lay = pdref.Layout()
gen = torch.Generator(device=dev); gen.manual_seed(5)
rew = torch.zeros(n, device=dev); done = torch.zeros(n, device=dev, dtype=torch.int32)
for t in range(pre + 1):
    if t % 33 == 0:
        a = (torch.rand((n, 2), device=dev, generator=gen) * 2 - 1).contiguous(); torch.cuda.synchronize()
    if t == pre:
        before = b.snapshot()
    b.env_step(a, 1.0 / 333.0, None, rew, done)
c = b.debug_warp_clocks()
d = c[4096 + n * 12:4096 + n * 16].reshape(n, 4)
cyc = d[:, 0].astype(np.float64)
print("k_collide per-car kcycles: p50 %.1f p90 %.1f p99 %.1f max %.1f" % tuple(np.percentile(cyc, [50, 90, 99, 100]) / 1e3))
for e in np.argsort(-cyc)[:12]:
    r = before[:, e]
    ncell, tr = d[e, 1] >> 32, d[e, 1] & 0xffffffff; wr, cand = d[e, 2] >> 32, d[e, 2] & 0xffffffff
    pos = [lay.get(r, "chassis.p" + k) for k in "xyz"]; v = np.linalg.norm([lay.get(r, "chassis.v" + k) for k in "xyz"])
    hullc = ((int(d[e, 3]) >> 8) & 0xfffffff) * 16 / 1e3; stagec = (int(d[e, 3]) >> 36) * 16 / 1e3
    print("env %d: %.0f kcycles (hull loops %.0f k, of which staging %.0f k), cells %d, track rounds %d, wall rounds %d, candidates %d, hit %d | pos %.1f %.1f %.1f v %.1f frame %d oot %d" % (e, cyc[e] / 1e3, hullc, stagec, ncell, tr, wr, cand, int(d[e, 3]) & 1, pos[0], pos[1], pos[2], v, lay.get(r, "car.physFrame"), lay.get(r, "car.outOfTrackFlag")))
# distribution over all cars (what a thread-per-car pre-filter could discard)
ncell = (d[:, 1] >> 32).astype(np.int64); tr = (d[:, 1] & 0xffffffff).astype(np.int64); wr = (d[:, 2] >> 32).astype(np.int64); cand = (d[:, 2] & 0xffffffff).astype(np.int64)
odd = np.array([lay.get(before[:, e], "car.physFrame") & 1 for e in range(n)])
print("odd-frame cars %d of %d; cells mean %.1f; track rounds: 0 -> %.1f%%, mean %.2f; wall rounds: 0 -> %.1f%%, mean %.2f; candidates: 0 -> %.1f%%, mean %.2f" % (
    odd.sum(), n, ncell.mean(), 100 * (tr == 0).mean(), tr.mean(), 100 * (wr == 0).mean(), wr.mean(), 100 * (cand == 0).mean(), cand.mean()))
for name, m in (("no rounds at all", (tr == 0) & (wr == 0)), ("track rounds only", (tr > 0) & (wr == 0)), ("wall rounds, no candidate", (wr > 0) & (cand == 0)), ("candidates", cand > 0)):
    if m.any(): print("  %-28s %5.1f%% of cars, mean %.1f kcycles, share of total cycles %.1f%%" % (name, 100 * m.mean(), cyc[m].mean() / 1e3, 100 * cyc[m].sum() / cyc.sum()))
